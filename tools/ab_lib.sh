for rep in 1 2; do
for lib in "" "build/libsdtf_hint.so"; do
  SDTF_LIB=$lib python bench.py --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=[$lib]', round(j['value'],3), round(j['unet_step_ms'],3), round(j['decode_ms_per_batch'],2), j['clocks']['sm_mhz'], round(j['roofline']['achieved']))"
done; done
SDTF_LIB=build/libsdtf_hint.so python tools/bench_kernels.py attn 2>&1 | grep -E "\"legacy\": false" | grep -E "self 64x64|cross 64x64|self 32x32" | cut -c1-160
