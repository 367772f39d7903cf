#!/bin/bash
# r02m: batch-1 regime: PDL A/B, per-shape trace table
mkdir -p gpurun_out
for spec in "b1:SDTF_PDL=0" "b1_pdl:SDTF_PDL=1" "b1:SDTF_PDL=0" "b1_pdl:SDTF_PDL=1"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env $envs python bench.py --batch 1 --steps 3 --warmup 2 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), 'e2e', round(j['e2e']['value'],3))"
done 2>&1 | tee gpurun_out/r02m_b1_pdl_ab.log
SDTF_TRACE=1 python bench.py --batch 1 --steps 1 --warmup 1 --denoise-steps 2 --skip-cpu-baseline --profile-only > gpurun_out/trace_b1.out 2> gpurun_out/trace_b1.log
python tools/trace_table.py gpurun_out/trace_b1.log > gpurun_out/r02m_trace_table_b1.md; head -50 gpurun_out/r02m_trace_table_b1.md
