#!/bin/bash
# r02c: flash VAE attention, ControlNet epilogue injection, new bench keys, batch-1 trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r02c_gpu_tests.log
tail -60 gpurun_out/r02c_gpu_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench_engine.json 2> gpurun_out/r02c_bench_engine.err
tail -3 gpurun_out/r02c_bench_engine.err; cat gpurun_out/r02c_bench_engine.json
python bench.py --batch 1 --steps 3 --warmup 2 --skip-cpu-baseline > gpurun_out/r02c_bench_b1.json 2> gpurun_out/r02c_bench_b1.err
tail -3 gpurun_out/r02c_bench_b1.err; cat gpurun_out/r02c_bench_b1.json
SDTF_TRACE=1 python bench.py --batch 1 --steps 1 --warmup 1 --denoise-steps 2 --skip-cpu-baseline --profile-only > gpurun_out/trace_b1.out 2> gpurun_out/trace_b1.log
python tools/trace_table.py gpurun_out/trace_b1.log > gpurun_out/r02c_trace_table_b1.md; head -60 gpurun_out/r02c_trace_table_b1.md
