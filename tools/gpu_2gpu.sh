#!/bin/bash
# 2 GPUs: the CFG-split equality test, tools/cfg_split_check.py output, and bench.py --gpus 2 (weak line with cfg_split self-check and strong_scaling)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -5 | tee gpurun_out/r02_2gpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/cfg_split_check.py 2>/dev/null | grep "^{" | tee gpurun_out/r02_cfg_split_check.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_engine_n2.json 2> gpurun_out/r02_bench_engine_n2.err
tail -3 gpurun_out/r02_bench_engine_n2.err; python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_engine_n2.json').read().strip().splitlines()[-1])
print('N=2 weak img/s', round(j['value'],2), 'e2e', round(j['e2e']['value'],2)); print('cfg_split', j.get('cfg_split')); print('strong', j.get('strong_scaling'))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong --no-cfg-split > gpurun_out/r02_bench_engine_n2_strong.json 2>/dev/null
python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_engine_n2_strong.json').read().strip().splitlines()[-1])
print('N=2 --scaling strong img/s', round(j['value'],2), j['scaling'], j['config']['parallelism'])"
