#!/bin/bash
# r02i: full text of the racecheck report on the CTA-pair GEMM; per-role wait profile + ncu of the HBM-bound 1x1 GEMM (320->320 +res @64x64)
mkdir -p gpurun_out
python -c "from minsdtf_b200 import build; print(build.build_test_gemm())"
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 3 build/test_gemm case conv3x3_16x16 2>&1 | head -60 | tee gpurun_out/r02i_racecheck_full.log
SDTF_GEMM_PROFILE=1 SDTF_GEMM_VERBOSE=1 build/test_gemm bench b16_linear_4096x320_res 2>&1 | tail -8 | tee gpurun_out/r02i_gemm_prof_1x1.log
SDTF_GEMM_PROFILE=1 SDTF_GEMM_VERBOSE=1 build/test_gemm bench b16_ff2_4096x1280 2>&1 | tail -8 | tee -a gpurun_out/r02i_gemm_prof_1x1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm3 -s 3 -c 1 -f -o gpurun_out/r02i_full_linear320 build/test_gemm bench b16_linear_4096x320_res > gpurun_out/r02i_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02i_full_linear320.ncu-rep 12 > gpurun_out/r02i_ncu_full_linear320.txt 2>&1; head -34 gpurun_out/r02i_ncu_full_linear320.txt
