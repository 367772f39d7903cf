#!/bin/bash
# r02b: folded upsample conv, packed-fp32 softmax exponentials (attn2h), cooperative GroupNorm launch
mkdir -p gpurun_out
python -c "from minsdtf_b200 import build; print(build.build_test_gemm())"
timeout 600 build/test_gemm 2>&1 | grep -E "upconv|FAIL|PASSED|ERROR|EXCEPTION" | tee gpurun_out/r02b_test_gemm.log
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02b_gpu_tests.log
tail -45 gpurun_out/r02b_gpu_tests.log
for pp in 0 2 3 4; do
  echo "== SDTF_ATTN_PP16=$pp"
  SDTF_ATTN_PP16=$pp python tools/bench_kernels.py attn 2>/dev/null | grep -E '"d": 40' | grep -v '"legacy": true' | grep -v '"Nk": 77'
done | tee gpurun_out/r02b_attn_pp16.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02b_bench_engine.json 2> gpurun_out/r02b_bench_engine.err
tail -2 gpurun_out/r02b_bench_engine.err; cat gpurun_out/r02b_bench_engine.json
VAR=SDTF_GN_COOP A=1 B=0 bash tools/ab_env.sh 2>&1 | tee gpurun_out/r02b_gn_coop_ab.log
VAR=SDTF_ATTN_PP16 A=3 B=0 bash tools/ab_env.sh 2>&1 | tee gpurun_out/r02b_pp16_ab.log
