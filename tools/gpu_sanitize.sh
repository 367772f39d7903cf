#!/bin/bash
# compute-sanitizer passes over small cases of every hand-written kernel.  Logs -> gpurun_out/.
mkdir -p gpurun_out
python -c "from minsdtf_b200 import build; print(build.build_test_gemm())"
SAN=/usr/local/cuda/bin/compute-sanitizer
export SDTF_ATTN_PERSIST=2  # several work items per CTA of the persistent attention kernel at sanitizer sizes
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  for cs in conv3x3_16x16 linear_geglu conv3x3_temb_res splitk_8x8_b2_silu upconv conv1x1_concat_slice linear_n1280_res; do
    echo "=== $tool test_gemm case $cs"
    timeout 600 $SAN --tool $tool --print-limit 20 build/test_gemm case $cs 2>&1 | grep -E "CASE|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" | head -12
  done
  echo "=== $tool tools/sanitize_cases.py"
  timeout 1500 $SAN --tool $tool --print-limit 20 python tools/sanitize_cases.py 2>&1 | grep -E " ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error|Error" | head -40
done 2>&1 | tee gpurun_out/r02_sanitizer.log
