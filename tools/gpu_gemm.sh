#!/bin/bash
# standalone conv/GEMM kernel check + micro-benchmark in the three dispatch modes
mkdir -p gpurun_out
for mode in default cg1 cg2; do
  case $mode in default) unset SDTF_GEMM_CG;; cg1) export SDTF_GEMM_CG=1;; cg2) export SDTF_GEMM_CG=2;; esac
  echo "=== mode $mode ===" | tee -a gpurun_out/gemm_test.log
  timeout 300 build/test_gemm bench 2>&1 | tee -a gpurun_out/gemm_test.log | grep -v "^CASE.*OK" | tail -40
done
