#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
python tools/bench_kernels.py attn 2>&1 | tee gpurun_out/bench_attn.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_engine.json 2> gpurun_out/bench_engine.err
tail -3 gpurun_out/bench_engine.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_engine.json'))
print({k:j[k] for k in ('value','ms_per_step','unet_step_ms','unet_step_tflops','decode_ms_per_batch','gpu_launches','clocks')}, j['e2e'], j['roofline'])
PY
