#!/bin/bash
# ncu --set full of the persistent GEMM kernel on representative shapes (standalone harness, 1 GPU)
mkdir -p gpurun_out
for c in ${CASES:-b16_linear_4096x320_res b16_geglu_4096 b16_conv3x3_64_320}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm3 -s 3 -c 1 -f -o gpurun_out/g3_$c build/test_gemm bench $c > gpurun_out/ncu_$c.log 2>&1
  tail -1 gpurun_out/ncu_$c.log
done
