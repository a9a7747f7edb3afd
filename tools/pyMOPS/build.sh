#!/usr/bin/env bash
# Builds tools/pyMOPS/pyMOPS<ext>.so (pybind11) against libmops_api.so / libmops_b200.so.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$HERE/../.."
PY="${PYTHON:-python}"
EXT="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
INC="$($PY -m pybind11 --includes)"
CUDA_LIB="${CUDA_HOME:-/usr/local/cuda}/lib64"
g++ -std=c++17 -O2 -fPIC -shared -fvisibility=hidden $INC -I"$ROOT/include" -o "$HERE/pyMOPS$EXT" "$HERE/bindings.cpp" \
    -L"$ROOT/mops_b200" -lmops_api -lmops_b200 -Wl,-rpath,"$ROOT/mops_b200" -Wl,-rpath,"$CUDA_LIB"
echo "built $HERE/pyMOPS$EXT"
