#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/gemm_prof.log
timeout 300 build/test_gemm 2>&1 | tail -3
for dbg in 0; do
for c in vae_b4_conv3x3_512_128 vae_b4_conv3x3_256_256 b16_conv3x3_64_320 b16_conv3x3_32_640 b16_linear_4096x320_res b16_geglu_4096; do
  for cg in 2; do
  echo "=== $c debug=$dbg cg=$cg" >> gpurun_out/gemm_prof.log
  SDTF_GEMM_CG=$cg SDTF_GEMM_DEBUG=$dbg SDTF_GEMM_PROFILE=1 timeout 120 build/test_gemm bench $c 2>&1 | grep "gemm3-prof" | tail -1 >> gpurun_out/gemm_prof.log
  done
done
done
MODES="default cg1 cg2 bn320" bash tools/gpu_gemm.sh > /dev/null 2>&1
