#!/usr/bin/env bash
# Builds mops_b200/libmops_b200.so (the C-ABI shared library) for sm_100a, in-tree.
#   -fmad=false : no FMA contraction -- the reference's CPU build rounds every product before
#                 the add and cell decisions must be bit-exact (see csrc/dmath.cuh)
#   -lineinfo   : ncu source page maps to these files
# Two translation units: engine.cu (the single-GPU engine: kernels + C ABI) and dist.cu (the multi-GPU layer; NCCL is
# loaded at run time, there is no link dependency).  Objects are cached under csrc/_build/<hash of the flags> and rebuilt
# when a source they include is newer.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${MOPS_OUT:-$HERE/../libmops_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-O2,-fno-fast-math"
KEY="$(echo "$FLAGS ${MOPS_DEFS:-}" | md5sum | cut -c1-12)"
OBJ="$HERE/_build/$KEY"
mkdir -p "$OBJ"
stale() { # stale <object> <sources...>
    local o="$1"; shift
    [ -f "$o" ] || return 0
    for s in "$@"; do [ "$s" -nt "$o" ] && return 0; done
    return 1
}
HDR="$HERE/../../include/mops_b200.h"
if [ -n "${MOPS_PTXAS_V:-}" ] || stale "$OBJ/engine.o" "$HERE/engine.cu" "$HERE"/*.cuh "$HDR"; then
    "$NVCC" $FLAGS ${MOPS_PTXAS_V:+-Xptxas -v} ${MOPS_DEFS:-} -c -o "$OBJ/engine.o" "$HERE/engine.cu" &
fi
if stale "$OBJ/dist.o" "$HERE/dist.cu" "$HDR"; then
    "$NVCC" $FLAGS -c -o "$OBJ/dist.o" "$HERE/dist.cu" &
fi
wait
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ/engine.o" "$OBJ/dist.o" -lcudart -lpthread -ldl
echo "built $OUT"
