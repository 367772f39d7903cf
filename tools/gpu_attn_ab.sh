#!/bin/bash
# attention kernel A/B on one box: parity of the attention cases, then tools/bench_kernels.py attn under each environment
# (arguments; "-" = default environment).  ATTN_AB_CASES selects the bench lines (regex, default: the two self-attention levels)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention or attn" 2>&1 | tail -5 | tee gpurun_out/attn_ab_tests.log
: > gpurun_out/attn_ab.log
cases=${ATTN_AB_CASES:-'"self (64x64|32x32)'}
for envs in "$@"; do
  echo "== $envs" | tee -a gpurun_out/attn_ab.log
  [ "$envs" = "-" ] && envs="SDTF_NOP=1"
  env $envs timeout 300 python tools/bench_kernels.py attn 2>&1 | grep -E "$cases" | grep '"legacy": false' | tee -a gpurun_out/attn_ab.log
done
