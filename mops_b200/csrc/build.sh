#!/usr/bin/env bash
# Builds mops_b200/libmops_b200.so (the C-ABI shared library) for sm_100a, in-tree.
#   -fmad=false : no FMA contraction -- the reference's CPU build rounds every product before
#                 the add and cell decisions must be bit-exact (see csrc/dmath.cuh)
#   -lineinfo   : ncu source page maps to these files
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${MOPS_OUT:-$HERE/../libmops_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
    -Xcompiler -fPIC,-O2,-fno-fast-math -shared ${MOPS_PTXAS_V:+-Xptxas -v} ${MOPS_DEFS:-} \
    -o "$OUT" "$HERE/engine.cu" -lcudart -lpthread
echo "built $OUT"
