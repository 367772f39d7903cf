#!/bin/bash
MODES="default cg1" bash tools/gpu_gemm.sh 2>&1 | grep -v "^BENCH" | tail -8
bash tools/gpu_check.sh
