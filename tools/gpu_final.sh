#!/bin/bash
# round-end verification: parity tests, both bench arms, ncu launch list and SDTF_TRACE table of the final code
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/gpu_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_engine.json 2> gpurun_out/bench_engine.err
tail -2 gpurun_out/bench_engine.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --denoise-steps 2 --no-graph --skip-cpu-baseline --profile-only > gpurun_out/ncu_bench.log 2>&1
SDTF_TRACE=1 python bench.py --steps 1 --warmup 1 --denoise-steps 2 --no-graph --skip-cpu-baseline --profile-only > gpurun_out/trace.out 2> gpurun_out/trace.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_engine.json'))
print({k:j[k] for k in ('value','ms_per_step','unet_step_ms','unet_step_frac_of_sustained_peak','decode_ms_per_batch','gpu_launches','clocks')}, j['e2e']['value'], j['roofline']['frac'], j['cpu_baseline']['value'])
print(open('gpurun_out/bench_reference.json').read()[:200])
PY
