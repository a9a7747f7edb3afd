#!/usr/bin/env bash
# Builds the engine variants that are queued for an A/B on B200 (scripts/gpu_ab.sh times every build_variants/*.so):
#   base          the committed defaults
#   roll_reloc    -DMOPS_ROLL_RELOC=1    rolled once-per-step cell relocation (smaller hot code)
#   cold_generic  -DMOPS_COLD_GENERIC=1  generic (non-hexagon) evaluation of 6-wide meshes out of line
#   roll_cold     both
#   sq_filter     -DMOPS_SQ_FILTER=1     streamline zero-velocity tests on the squared norm (time it with scripts/bench_secondary.py: C3)
# Parity for a variant: MOPS_B200_LIB=$PWD/build_variants/<name>.so python -m pytest tests -m gpu -q
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p build_variants
b() { name=$1; shift; MOPS_OUT=$PWD/build_variants/$name.so MOPS_DEFS="$*" bash mops_b200/csrc/build.sh; }
b base &
b roll_reloc -DMOPS_ROLL_RELOC=1 &
b cold_generic -DMOPS_COLD_GENERIC=1 &
b roll_cold -DMOPS_ROLL_RELOC=1 -DMOPS_COLD_GENERIC=1 &
b sq_filter -DMOPS_SQ_FILTER=1 &
wait
