#!/bin/bash
# GPU visit: the whole -m gpu suite (with durations), then the engine bench line and the per-operator trace table.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/gpu_tests.log
tail -40 gpurun_out/gpu_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_engine.json 2> gpurun_out/bench_engine.err
tail -3 gpurun_out/bench_engine.err
cat gpurun_out/bench_engine.json
SDTF_TRACE=1 python bench.py --steps 1 --warmup 1 --denoise-steps 2 --no-graph --skip-cpu-baseline --profile-only > gpurun_out/trace.out 2> gpurun_out/trace.log
python tools/trace_table.py gpurun_out/trace.log > gpurun_out/trace_table.md 2>/dev/null; head -40 gpurun_out/trace_table.md
