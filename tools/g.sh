#!/bin/bash
# usage: tools/g.sh <timeout_s> '<command>'   — rebuild every native artefact, then run the command on a B200 via gpurun
set -e
cd "$(dirname "$0")/.."
python -c "
import sys; sys.path.insert(0,'.')
from minsdtf_b200 import build
build.build_lib(); build.build_test_gemm()"
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
