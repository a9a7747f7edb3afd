# Round 2, call 5 (1 GPU): A/B of the reduced-load fast path (computed edge vectors, 24-byte velocity loads) at 3/4/5 resident
# blocks per SM, then source-level ncu captures of the 3- and 4-block builds on the reduced-footprint launch
set -x
mkdir -p gpurun_out
for f in build_variants/*.so; do
  MOPS_B200_LIB=$PWD/$f timeout 200 python bench.py --level 8 --particles 16000000 --interval-steps 60 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ab.log 2> gpurun_out/ab.err || { echo "$f FAILED"; tail -3 gpurun_out/ab.err; continue; }
  tail -1 gpurun_out/ab.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],2), 'value', round(d['value']/1e9,4), 'exec_frac', round(d['config']['executed_fraction'],4))"
done 2>&1 | tee gpurun_out/r02_ab_fast2.txt
CMD="python bench.py --level 7 --particles 4000000 --interval-steps 30 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
for v in fast2 fast2_m4; do
  MOPS_B200_LIB=$PWD/build_variants/$v.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_advect -s 3 -c 1 -f -o gpurun_out/r02_prof_$v $CMD > gpurun_out/ncu_$v.log 2>&1
  tail -2 gpurun_out/ncu_$v.log | cut -c1-200
done
( MOPS_B200_LIB=$PWD/build_variants/fast2.so python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5 | tee gpurun_out/r02_pytest_fast2.txt
