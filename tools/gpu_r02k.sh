#!/bin/bash
# r02k: one GEMM instantiation per epilogue variant: tests + A/B (fold all three LayerNorms vs two)
mkdir -p gpurun_out
python -c "from minsdtf_b200 import build; print(build.build_test_gemm())"
timeout 600 build/test_gemm 2>&1 | grep -E "FAIL|PASSED|ERROR|EXCEPTION" | tee gpurun_out/r02k_test_gemm.log
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r02k_gpu_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r02k_gpu_tests.log
bash tools/ab_env3.sh "modes_fold2:SDTF_LN_FOLD=2" "modes_fold_all:SDTF_LN_FOLD=1" "general_fold2:SDTF_GEMM_LEAN=0" 2>&1 | tee gpurun_out/r02k_ab.log
for spec in "b1_modes:SDTF_GEMM_LEAN=1" "b1_general:SDTF_GEMM_LEAN=0" "b1_fold_all:SDTF_LN_FOLD=1"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env $envs python bench.py --batch 1 --steps 3 --warmup 2 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), {k: round(v['ms'],3) for k,v in j['operator_classes']['denoise_step'].items()})"
done 2>&1 | tee -a gpurun_out/r02k_ab.log
