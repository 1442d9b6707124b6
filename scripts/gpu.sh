#!/bin/bash
# build in-tree, stop if the build fails, then run the given command on a B200 through gpurun
# usage: scripts/gpu.sh <timeout seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "
from trlda_b200 import build
build.build_all(verbose=False)" 2>&1 | grep -E "error|Error" && { echo BUILD FAILED; exit 1; }
python -c "from trlda_b200 import capi; capi.lib()" || { echo LOAD FAILED; exit 1; }
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
