#!/usr/bin/env bash
# Builds the C++ drop-in library (mops_b200/libmops_api.so: include/api/MOPS.h over the C ABI)
# and the three tutorials against it.  Needs mops_b200/libmops_b200.so (mops_b200/csrc/build.sh).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$HERE/.."
CXX="${CXX:-g++}"
CUDA_LIB="${CUDA_HOME:-/usr/local/cuda}/lib64"
$CXX -std=c++17 -O2 -fPIC -shared -I"$ROOT/include" -o "$ROOT/mops_b200/libmops_api.so" "$ROOT/mops_b200/host/mops_api.cpp" "$ROOT/mops_b200/host/mops_reader.cpp" "$ROOT/mops_b200/host/mpas_io.cpp" -I"$ROOT/mops_b200/host" \
    -L"$ROOT/mops_b200" -lmops_b200 -pthread -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$CUDA_LIB"
mkdir -p "$HERE/bin"
for t in streamLine pathLine reMapping; do
    $CXX -std=c++17 -O2 -I"$ROOT/include" -I"$HERE" -o "$HERE/bin/$t" "$HERE/$t.cpp" \
        -L"$ROOT/mops_b200" -lmops_api -lmops_b200 -Wl,-rpath,"$ROOT/mops_b200" -Wl,-rpath,"$CUDA_LIB"
done
$CXX -std=c++17 -O2 -I"$ROOT/include" -I"$ROOT/mops_b200/host" -o "$HERE/bin/mops_cli" "$ROOT/CLI/main.cpp" \
    -L"$ROOT/mops_b200" -lmops_api -lmops_b200 -Wl,-rpath,"$ROOT/mops_b200" -Wl,-rpath,"$CUDA_LIB"
echo "built $ROOT/mops_b200/libmops_api.so and $HERE/bin/{streamLine,pathLine,reMapping,mops_cli}"
