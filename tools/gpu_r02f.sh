#!/bin/bash
# r02f: per-pass LayerNorm slots (batch invariance restored), dense heads default; all-new vs all-old switches on one box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r02f_gpu_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r02f_gpu_tests.log
bash tools/ab_env3.sh "new:SDTF_LN_FOLD=1" "fold12_only:SDTF_LN_FOLD=2" "old_switches:SDTF_LN_FOLD=0 SDTF_ATTN_PP16=0 SDTF_HEAD_DENSE=0 SDTF_HEAD_CLIP=0" 2>&1 | tee gpurun_out/r02f_ab.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02f_bench_engine.json 2> gpurun_out/r02f_bench_engine.err
tail -2 gpurun_out/r02f_bench_engine.err; cut -c1-600 gpurun_out/r02f_bench_engine.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; cat gpurun_out/r02f_bench_reference.json
