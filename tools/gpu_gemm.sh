#!/bin/bash
# standalone conv/GEMM kernel check + micro-benchmark in several dispatch modes
mkdir -p gpurun_out
rm -f gpurun_out/gemm_test.log
for mode in ${MODES:-default cg1 cg2 bn320 bn512 bn160}; do
  unset SDTF_GEMM_CG SDTF_GEMM_BN
  case $mode in cg1) export SDTF_GEMM_CG=1;; cg2) export SDTF_GEMM_CG=2;; bn320) export SDTF_GEMM_BN=320;; bn512) export SDTF_GEMM_BN=512;; bn160) export SDTF_GEMM_BN=160;; esac
  echo "=== mode $mode ===" | tee -a gpurun_out/gemm_test.log
  timeout 300 build/test_gemm bench 2>&1 | tee -a gpurun_out/gemm_test.log | grep -v "^CASE.*OK" | tail -40
done
