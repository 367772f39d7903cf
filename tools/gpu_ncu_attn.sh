#!/bin/bash
# ncu --set full of the 64x64 self-attention kernel under each environment given ("name:ENV=..." arguments)
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 600 ncu --set full --clock-control none --import-source on -k regex:${ATTN_NCU_RX:-attn2h} -s ${ATTN_NCU_SKIP:-2} -c 1 -f -o gpurun_out/attn_$name \
    python tools/bench_kernels.py attn > gpurun_out/ncu_attn_$name.log 2>&1
  tail -1 gpurun_out/ncu_attn_$name.log
  python tools/ncu_summary.py gpurun_out/attn_$name.ncu-rep 30 > gpurun_out/ncu_attn_$name.txt 2>&1; head -24 gpurun_out/ncu_attn_$name.txt
done
