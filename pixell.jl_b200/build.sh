#!/bin/bash
# Builds pixell.jl_b200/lib/libpixsht.so (the CUDA engine + C ABI + libsharp2-compatible shim) for sm_100a, in tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$HERE/lib"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC,-O3 \
    -ccbin /usr/bin/g++ ${PIXSHT_NVCC_EXTRA:-} \
    -o "$HERE/lib/libpixsht.so" "$HERE/csrc/pixsht.cu" "$HERE/csrc/sharp_shim.cu"
echo "built $HERE/lib/libpixsht.so"
