#!/bin/bash
# TEST INFRASTRUCTURE ONLY: compiles the kernels of pixell.jl_b200/csrc for the host with tests/emu/cuda_emu.h so that
# their logic can be exercised in the GPU-less build container.  Never loaded by the product.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
mkdir -p "$HERE/_build"
/usr/bin/g++ -O2 -g -std=c++17 -fPIC -shared -pthread -x c++ -DPIXSHT_EMU -include "$HERE/cuda_emu.h" \
    -Wall -Wno-unknown-pragmas -Wno-unused-function -Wno-unused-variable \
    -o "$HERE/_build/libpixsht_emu.so" "$ROOT/pixell.jl_b200/csrc/pixsht.cu" "$ROOT/pixell.jl_b200/csrc/sharp_shim.cu"
echo "built $HERE/_build/libpixsht_emu.so"
