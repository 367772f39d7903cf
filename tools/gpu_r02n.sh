#!/bin/bash
# r02n: GroupNorm partial-sum chains A/B (batch 16 and batch 1), norm unit tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02n_tests.log
bash tools/ab_env3.sh "chains4:SDTF_GN_CHAINS=4" "chains1:SDTF_GN_CHAINS=1" 2>&1 | tee gpurun_out/r02n_gn_chains_ab.log
for spec in "b1_chains4:SDTF_GN_CHAINS=4" "b1_chains1:SDTF_GN_CHAINS=1" "b1_chains4:SDTF_GN_CHAINS=4" "b1_chains1:SDTF_GN_CHAINS=1"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env $envs python bench.py --batch 1 --steps 3 --warmup 2 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), {k: round(v['ms'],3) for k,v in j['operator_classes']['denoise_step'].items()})"
done 2>&1 | tee -a gpurun_out/r02n_gn_chains_ab.log
