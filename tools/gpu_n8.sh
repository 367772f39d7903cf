#!/bin/bash
# usage: gpurun --gpus 8 -- 'bash tools/gpu_n8.sh': the engine arm of bench.py at 8 GPUs as the driver launches it (weak line +
# cfg_split self-check + strong arm); the reference arm is CPU-only on rank 0 and is exercised at N = 2 / 4 (gpu_ngpu.sh)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r02_bench_engine_n8.json 2> gpurun_out/r02_bench_engine_n8.err
tail -3 gpurun_out/r02_bench_engine_n8.err; python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_engine_n8.json').read().strip().splitlines()[-1])
print('N=8 weak img/s', round(j['value'],2), 'e2e', round(j['e2e']['value'],2), 'clocks', j['clocks']); print('cfg_split', j.get('cfg_split')); print('strong', j.get('strong_scaling'))"
