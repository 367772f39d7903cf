#!/bin/bash
mkdir -p gpurun_out
python tools/ablate.py 2>&1 | tail -2 | tee gpurun_out/ablate.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2q -s 2 -c 1 -f -o gpurun_out/attn2q python tools/bench_kernels.py attn > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
