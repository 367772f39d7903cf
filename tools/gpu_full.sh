#!/bin/bash
# One full GPU visit: parity tests, both bench arms, ncu launch list of a 2-step generation, and ncu --set full captures
# of the dominant kernels (implicit-GEMM conv, d=40 self-attention, one-kernel GroupNorm, cross-attention) and the
# SDTF_TRACE operator table.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/gpu_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_engine.json 2> gpurun_out/bench_engine.err
tail -3 gpurun_out/bench_engine.err
cat gpurun_out/bench_engine.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json
python tools/bench_kernels.py all 2>&1 | tee gpurun_out/bench_kernels.jsonl
python tools/ablate.py 2>&1 | tail -1 | tee gpurun_out/ablation.json
# launch list: 2 denoise steps, eager launches (no graph) so each kernel is one ncu row
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --denoise-steps 2 --no-graph --skip-cpu-baseline --profile-only > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
wc -l gpurun_out/launches.csv
# full captures (one launch each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm3 -s 3 -c 1 -f -o gpurun_out/full_conv3x3_64_320 \
  build/test_gemm bench b16_conv3x3_64_320 > gpurun_out/ncu_full_conv.log 2>&1
tail -1 gpurun_out/ncu_full_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2 -s 2 -c 1 -f -o gpurun_out/full_attn_self64 \
  python tools/bench_kernels.py attn > gpurun_out/ncu_full_attn.log 2>&1
tail -1 gpurun_out/ncu_full_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_fused -s 8 -c 1 -f -o gpurun_out/full_gn_fused \
  python bench.py --steps 1 --warmup 0 --denoise-steps 1 --no-graph --skip-cpu-baseline --profile-only > gpurun_out/ncu_full_gn.log 2>&1
tail -1 gpurun_out/ncu_full_gn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn -s 1 -c 1 -f -o gpurun_out/full_xattn \
  python tools/bench_kernels.py attn > gpurun_out/ncu_full_xattn.log 2>&1
tail -1 gpurun_out/ncu_full_xattn.log
SDTF_TRACE=1 python bench.py --steps 1 --warmup 1 --denoise-steps 2 --no-graph --skip-cpu-baseline --profile-only > gpurun_out/trace.out 2> gpurun_out/trace.log
ls -la gpurun_out
