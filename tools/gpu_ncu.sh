#!/bin/bash
# ncu evidence for round 2: launch list (time + DRAM bytes) of one eager denoise step + decode at the benchmark batch, and
# --set full captures of the dominant kernels.  A number printed by a run under ncu is never a bench value.
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 0 --denoise-steps 1 --no-graph --skip-cpu-baseline --profile-only \
  > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log; wc -l gpurun_out/r02_launches.csv
python tools/summarize_launches.py gpurun_out/r02_launches.csv gpurun_out/r02_step_traffic.json > gpurun_out/r02_launches_summary.md; head -40 gpurun_out/r02_launches_summary.md
full() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/r02_full_$name "$@" > gpurun_out/r02_ncu_full_$name.log 2>&1
  tail -1 gpurun_out/r02_ncu_full_$name.log
  python tools/ncu_summary.py gpurun_out/r02_full_$name.ncu-rep 12 > gpurun_out/r02_ncu_full_$name.txt 2>&1; head -30 gpurun_out/r02_ncu_full_$name.txt
}
full attn2h attn2h 2 python tools/bench_kernels.py attn
full conv3x3_64_320 conv_gemm3 3 build/test_gemm bench b16_conv3x3_64_320
full vattn vattn 0 python bench.py --steps 1 --warmup 0 --denoise-steps 1 --no-graph --skip-cpu-baseline --profile-only
full gn_fused gn_fused 8 python bench.py --steps 1 --warmup 0 --denoise-steps 1 --no-graph --skip-cpu-baseline --profile-only
rm -f gpurun_out/*.ncu-rep.tmp
ls -la gpurun_out | tail -20
