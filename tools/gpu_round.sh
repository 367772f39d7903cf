#!/bin/bash
# One GPU visit: bench (engine + reference arm), then an ncu launch list of one short generation.
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_engine.json 2> gpurun_out/bench_engine.err
tail -3 gpurun_out/bench_engine.err
cat gpurun_out/bench_engine.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json
# launch list: 2 denoise steps, eager launches (no graph) so each kernel is one ncu row
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --denoise-steps 2 --no-graph --skip-cpu-baseline --profile-only > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
wc -l gpurun_out/launches.csv
