#!/bin/bash
# parity tests, then an A/B of one environment switch on the same box: bench.py with $AB_VAR=$AB_A and $AB_VAR=$AB_B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
for v in $AB_A $AB_B $AB_A $AB_B; do
  env $AB_VAR=$v timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_${AB_VAR}_$v.json 2> gpurun_out/bench_engine.err
  tail -3 gpurun_out/bench_engine.err
  python - <<PY
import json
j=json.load(open('gpurun_out/bench_${AB_VAR}_$v.json'))
print('$AB_VAR=$v', {k:j[k] for k in ('value','unet_step_ms','decode_ms_per_batch','gpu_launches','clocks')}, j['e2e']['value'], j['roofline']['achieved'])
PY
done
