#!/bin/bash
# r02d: head-clipped stores (padded head stride) vs dense heads, small-batch schedules (split-K classes, xattn z-split, GN partitions)
mkdir -p gpurun_out
python -c "from minsdtf_b200 import build; print(build.build_test_gemm())"
timeout 600 build/test_gemm 2>&1 | grep -E "FAIL|PASSED|ERROR|EXCEPTION|invariance" | tee gpurun_out/r02d_test_gemm.log
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r02d_gpu_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r02d_gpu_tests.log
VAR=SDTF_HEAD_DENSE A=0 B=1 bash tools/ab_env.sh 2>&1 | tee gpurun_out/r02d_head_dense_ab.log
b1() { python bench.py --batch 1 --steps 3 --warmup 2 --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3), 'decode', round(j['decode_ms_per_batch'],2), {k: round(v['ms'],3) for k,v in j['operator_classes']['denoise_step'].items()})"; }
echo "== batch 1 (UNet batch 2)" | tee gpurun_out/r02d_batch1_ab.log
(b1 default; SDTF_SPLITK_SMALL=0 b1 splitk_small_off; SDTF_GN_SMALL=0 b1 gn_small_off; SDTF_SPLITK_SMALL=0 SDTF_GN_SMALL=0 b1 both_off; b1 default) 2>&1 | tee -a gpurun_out/r02d_batch1_ab.log
python bench.py --batch 2 --steps 3 --warmup 2 --skip-cpu-baseline 2>/dev/null > gpurun_out/r02d_bench_b2.json; python -c "
import json; j=json.loads(open('gpurun_out/r02d_bench_b2.json').read().strip().splitlines()[-1]); print('batch2 img/s', round(j['value'],3), 'step_ms', round(j['unet_step_ms'],3))"
