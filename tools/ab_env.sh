#!/bin/bash
# usage: VAR=SDTF_DUAL A=0 B=1 bash tools/ab_env.sh  — bench.py alternately under VAR=A / VAR=B on the same box (2 rounds)
for rep in 1 2; do
for v in $A $B; do
  env $VAR=$v python bench.py --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', 'img/s', round(j['value'],3), 'e2e', round(j['e2e']['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), 'MHz', j['clocks']['sm_mhz'])"
done; done
