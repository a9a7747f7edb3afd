# Reduced-footprint ncu capture of the advection kernel (the full-size launch keeps ~70 GB resident, which ncu
# saves/restores around every replay pass), then an A/B of occupancy variants (build_variants/*.so).
set -x
CMD="python bench.py --level 7 --particles 4000000 --interval-steps 30 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_small.log 2> gpurun_out/plain_small.err && tail -1 gpurun_out/plain_small.log | cut -c1-200 &&
timeout 330 ncu --set full --clock-control none --import-source on -k regex:k_advect -s 3 -c 1 -f -o gpurun_out/prof_advect_small $CMD > gpurun_out/ncu_full_small.log 2>&1
tail -2 gpurun_out/ncu_full_small.log | cut -c1-200
for f in build_variants/base.so build_variants/b64m7.so build_variants/b128m4.so; do
  MOPS_B200_LIB=$PWD/$f timeout 200 python bench.py --level 8 --particles 16000000 --interval-steps 60 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab.log 2> gpurun_out/ab.err || { echo "$f FAILED"; tail -3 gpurun_out/ab.err; continue; }
  tail -1 gpurun_out/ab.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],2), 'value', round(d['value']/1e9,4))"
done 2>&1 | tee gpurun_out/ab_summary.txt
ls -la gpurun_out/
