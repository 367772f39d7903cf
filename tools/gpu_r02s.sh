#!/bin/bash
mkdir -p gpurun_out
timeout 600 build/test_gemm 2>&1 | grep -E "FAIL|PASSED|ERROR|EXCEPTION" | tee gpurun_out/r02s_test_gemm.log
build/test_gemm bench b16_ 2>&1 | grep BENCH | tee gpurun_out/r02s_bench_cases.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_golden.py -m gpu -q -x 2>&1 | tail -3
bash tools/gpu_bench1.sh 2>&1 | head -8
