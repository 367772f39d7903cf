for rep in 1 2; do
for cfg in "SDTF_ATTN_DEFER=0" "SDTF_ATTN_DEFER=16" "SDTF_SPLITK=0" "SDTF_XATTN=0"; do
  env $cfg python bench.py --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(j['value'],3), round(j['unet_step_ms'],3), j['clocks']['sm_mhz'])"
done; done
