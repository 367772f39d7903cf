#!/bin/bash
# r02p: trimmed GEMM code (fast SiLU division, multiply-high tile decode): correctness + isolated shapes + step
mkdir -p gpurun_out
timeout 600 build/test_gemm 2>&1 | grep -E "FAIL|PASSED|ERROR|EXCEPTION" | tee gpurun_out/r02p_test_gemm.log
build/test_gemm bench b16_ 2>&1 | grep BENCH | tee gpurun_out/r02p_bench_cases.log
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r02p_gpu_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r02p_gpu_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02p_bench_engine.json 2> gpurun_out/r02p_bench_engine.err
python -c "
import json; j=json.loads(open('gpurun_out/r02p_bench_engine.json').read().strip().splitlines()[-1])
print('img/s', round(j['value'],3), 'e2e', round(j['e2e']['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), 'MHz', j['clocks']['sm_mhz'], {k: round(v['ms'],3) for k,v in j['operator_classes']['denoise_step'].items()}, 'roofline', round(j['roofline']['frac'],3), round(j['roofline']['frac_sustained'],3))"
python bench.py --batch 1 --steps 3 --warmup 2 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b1', 'img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2))"
