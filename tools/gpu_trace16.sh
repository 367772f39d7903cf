#!/bin/bash
mkdir -p gpurun_out
SDTF_TRACE=1 python bench.py --steps 1 --warmup 1 --denoise-steps 2 --skip-cpu-baseline --profile-only > gpurun_out/trace.out 2> gpurun_out/trace.log
python tools/trace_table.py gpurun_out/trace.log > gpurun_out/r02q_trace_table.md; head -75 gpurun_out/r02q_trace_table.md
