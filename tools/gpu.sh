#!/bin/bash
# Rebuild the CUDA library in tree, then hand the command to gpurun (the built .so travels with the snapshot).
set -e
cd "$(dirname "$0")/.."
bash pixell.jl_b200/build.sh >/dev/null
T="${GPU_TIMEOUT:-1200}"
exec /usr/local/graft/bin/gpurun --timeout "$T" ${GPU_N:+--gpus $GPU_N} -- "$@"
