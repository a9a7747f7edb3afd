#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the reference's own TBB/CPU implementation of the hot path, UNMODIFIED, from the
# sources where they lie under $MOPS_REFERENCE (default /root/reference), plus our flat-C
# driver (oracle/ref_driver.cpp), into the git-ignored oracle/_ref/:
#     oracle/_ref/libmops_ref.so      -O2, no -march, no -ffast-math  (parity oracle + CPU baseline)
# The reference's own build system (CMake + TBB + netCDF + ndarray + yaml-cpp + VTK) is not
# used: none of those dependencies exist in this image and none is needed by the hot path.
# The four headers they would provide are stubbed in oracle/shim/ (see each file's header).
# No reference source is copied into the repo.
#
# tbb::parallel_for is backed by OpenMP (schedule(dynamic,64)); every call site iterates
# over independent particles / pixels / (vertex,level) pairs, so results do not depend on the
# thread count (tests/test_oracle_ref.py checks 1 thread == N threads bit-for-bit).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${MOPS_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
SHIM="$HERE/shim"
if [ ! -d "$R/src" ]; then
    echo "build_ref.sh: reference tree not found at $R (expected on the GPU box: prebuilt oracle/_ref is used)" >&2
    exit 3
fi
mkdir -p "$OUT/obj"
DEFS="-DMOPS_USE_CPU=1 -DMOPS_USE_GPU=0 -DMOPS_USE_SYCL=0 -DMOPS_USE_CUDA=0 -DMOPS_USE_HIP=0 -DMOPS_USE_TBB=1 \
      -DMOPS_VTK=0 -DMOPS_MPI=0 -D_DEBUG=0 -DMOPS_ENABLE_TIMING=1 -DMOPS_SHIM_OMP"
CXXFLAGS="-std=c++17 -O2 -fPIC -fopenmp -w"
INCS="-I$SHIM -I$R/include -I$R/src"
TUS="Core/MOPS Core/MOPSApp Core/MPASOField Core/MPASOGrid Core/MPASOSolution Core/MPASOVisualizer \
     Common/MOPSFactory CPU/Common/CPUFactory GPU/Common/GPUFactory \
     CPU/TBB/MPASOSolutionTBB CPU/TBB/MPASOVisualizerTBB CPU/TBB/Kernel/MPASOVisualizerKernels CPU/TBB/Kernel/TBBKernel \
     IO/MPASOReader Utils/KDTree Utils/impl"
pids=()
for tu in $TUS; do
    obj="$OUT/obj/$(echo "$tu" | tr / _).o"
    extra=""
    # the one deliberate deviation: by-reference SetPixel for the TBB remap (see shim/setpixel_fix.h)
    if [ "$tu" = "CPU/TBB/Kernel/MPASOVisualizerKernels" ]; then extra="-include $SHIM/setpixel_fix.h"; fi
    g++ $CXXFLAGS $DEFS $INCS $extra -c "$R/src/$tu.cpp" -o "$obj" &
    pids+=($!)
done
g++ $CXXFLAGS $DEFS $INCS -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -fopenmp -o "$OUT/libmops_ref.so" "$OUT"/obj/*.o -lpthread
rm -rf "$OUT/obj"
echo "built $OUT/libmops_ref.so"
