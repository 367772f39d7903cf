#!/bin/bash
# usage: bash tools/ab_env3.sh "NAME1:ENV1=V ENV2=V" "NAME2:..." ...  — bench.py alternately under each environment on the same box
# (2 rounds), one line per run with the per-class operator times of a denoise step
for rep in 1 2; do
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env $envs python bench.py --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), 'MHz', j['clocks']['sm_mhz'], {k: round(v['ms'],3) for k,v in j['operator_classes']['denoise_step'].items()})"
done; done
