#!/bin/bash
# round-end verification of the committed code: the whole -m gpu suite, smoke(), both bench arms, batch-1 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/final_gpu_tests.log
grep -E "passed|failed|FAILED" gpurun_out/final_gpu_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; cut -c1-300 gpurun_out/final_bench_reference.json
python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench_engine.json 2> gpurun_out/final_bench_engine.err
tail -2 gpurun_out/final_bench_engine.err; cat gpurun_out/final_bench_engine.json
python bench.py --batch 1 --steps 3 --warmup 2 --skip-cpu-baseline > gpurun_out/final_bench_b1.json 2>/dev/null; python -c "
import json; j=json.loads(open('gpurun_out/final_bench_b1.json').read().strip().splitlines()[-1])
print('b1 img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), j['roofline']['frac'])"
