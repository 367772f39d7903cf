#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_ngpu.sh N': bench.py at N GPUs as the driver launches it (weak line + cfg_split self-check + strong arm)
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_engine_n$N.json 2> gpurun_out/r02_bench_engine_n$N.err
tail -3 gpurun_out/r02_bench_engine_n$N.err; python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_engine_n$N.json').read().strip().splitlines()[-1])
print('N=$N weak img/s', round(j['value'],2), 'e2e', round(j['e2e']['value'],2), 'clocks', j['clocks']); print('cfg_split', j.get('cfg_split')); print('strong', j.get('strong_scaling'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>/dev/null | cut -c1-200
