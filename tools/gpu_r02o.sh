#!/bin/bash
# r02o: memcheck over a whole tiny job; the five BASELINE configurations at full size through the public API
mkdir -p gpurun_out
timeout 1500 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_cases.py pipeline 2>&1 | grep -E " ok|ERROR SUMMARY|Invalid|error|Error" | head -30 | tee gpurun_out/r02o_memcheck_pipeline.log
timeout 1200 python tools/config_sweep.py 2>/dev/null | tee gpurun_out/r02o_config_sweep.jsonl
