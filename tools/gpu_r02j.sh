#!/bin/bash
# r02j: lean GEMM instantiation A/B
mkdir -p gpurun_out
python -c "from minsdtf_b200 import build; print(build.build_test_gemm())"
timeout 600 build/test_gemm 2>&1 | grep -E "FAIL|PASSED|ERROR|EXCEPTION" | tee gpurun_out/r02j_test_gemm.log
for v in 1 0; do echo "== SDTF_GEMM_LEAN=$v"; SDTF_GEMM_LEAN=$v build/test_gemm bench b16_ 2>&1 | grep BENCH; done | tee gpurun_out/r02j_lean_bench_cases.log
bash tools/ab_env3.sh "lean:SDTF_GEMM_LEAN=1" "general:SDTF_GEMM_LEAN=0" 2>&1 | tee gpurun_out/r02j_lean_ab.log
