# Round 2, call 2 (1 GPU): parity of the hexagon-fast-path engine (edge vectors in the record, unhalved areas, no-w instantiation,
# device-side live count, double-buffered HOST staging) + A/B against the round-1 engine at 3 / 4 / 5 resident blocks per SM.
set -x
mkdir -p gpurun_out
( MOPS_B200_LIB=$PWD/build_variants/hex2.so python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 | tee gpurun_out/r02_pytest_hex2.txt
bash scripts/gpu_ab.sh
cp gpurun_out/ab_summary.txt gpurun_out/r02_ab_hex2.txt
