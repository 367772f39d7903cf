#!/bin/bash
# r02e: LayerNorm folded into the GEMMs (parity + A/B), head layout A/B on one box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r02e_gpu_tests.log
grep -E "passed|failed|FAILED|rel err|psnr" gpurun_out/r02e_gpu_tests.log | head -60
bash tools/ab_env3.sh "fold+pad_clip:SDTF_LN_FOLD=1" "nofold+pad_clip:SDTF_LN_FOLD=0" "fold+dense:SDTF_HEAD_DENSE=1" "fold+pad_noclip:SDTF_HEAD_CLIP=0" 2>&1 | tee gpurun_out/r02e_ab.log
echo "== batch 1" | tee -a gpurun_out/r02e_ab.log
for spec in "fold:SDTF_LN_FOLD=1" "nofold:SDTF_LN_FOLD=0" "fold:SDTF_LN_FOLD=1"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env $envs python bench.py --batch 1 --steps 3 --warmup 2 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), {k: round(v['ms'],3) for k,v in j['operator_classes']['denoise_step'].items()})"
done 2>&1 | tee -a gpurun_out/r02e_ab.log
