#!/usr/bin/env bash
# Builds engine variants for an A/B on B200 (scripts/gpu_ab.sh times every build_variants/*.so on the level-8 / 16 M-particle /
# 60-step case).  Knobs: MOPS_ADV_BLOCK, MOPS_ADV_MINB, MOPS_FAST_UNROLL_SNAP, MOPS_FAST_LOAD24,
# MOPS_FAST_LDG.  Parity for a variant: MOPS_B200_LIB=$PWD/build_variants/<name>.so python -m pytest tests -m gpu -q
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p build_variants
b() { name=$1; shift; MOPS_OUT=$PWD/build_variants/$name.so MOPS_DEFS="$*" bash mops_b200/csrc/build.sh; }
b base &
b minb4 -DMOPS_ADV_MINB=4 &
b unroll_snap -DMOPS_FAST_UNROLL_SNAP=2 &
b load32 -DMOPS_FAST_LOAD24=0 &
wait
