#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -3 gpurun_out/bench1.err
python -c "
import json; j=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1])
print('img/s', round(j['value'],3), 'e2e', round(j['e2e']['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), 'MHz', j['clocks']['sm_mhz'])
print('roofline', {k:(round(v,3) if isinstance(v,float) else v) for k,v in j['roofline'].items() if k!='how' and k!='kernel'})
for part in ('denoise_step','vae_decode'): print(part, {k:{a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()} for k,v in j['operator_classes'][part].items()})"
python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --no-graph 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-graph: step_ms', round(j['unet_step_ms'],3), 'roofline', round(j['roofline']['achieved'],1), j['roofline']['kernel'])"
