#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -5
python tools/bench_kernels.py attn 2>&1 | grep -v "legacy\": true" | cut -c1-200
echo "--- previous full-row kernel"
SDTF_ATTN_2Q=1 python tools/bench_kernels.py attn 2>&1 | grep "d\": 40" | grep -v "legacy\": true" | cut -c1-200
