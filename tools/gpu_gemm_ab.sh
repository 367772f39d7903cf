#!/bin/bash
# A/B of two builds of the standalone GEMM harness in the same run (same box, same clocks)
mkdir -p gpurun_out
rm -f gpurun_out/gemm_ab.log
for rep in 1 2; do
for exe in build/test_gemm_old build/test_gemm; do
  echo "=== $exe rep $rep ===" | tee -a gpurun_out/gemm_ab.log
  SDTF_GEMM_VERBOSE=$([ $rep = 1 ] && echo 1 || echo 0) timeout 300 $exe bench b 2>&1 | grep -v "^CASE" | tee -a gpurun_out/gemm_ab.log | grep -c BENCH
done
done
