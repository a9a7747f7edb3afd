# A/B timing of differently built engine libraries (build_variants/*.so) on a small C5-shaped case (level-8 mesh, 16 M particles)
for f in build_variants/*.so; do
  MOPS_B200_LIB=$PWD/$f timeout 200 python bench.py --level 8 --particles 16000000 --interval-steps 60 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ab.log 2> gpurun_out/ab.err || { echo "$f FAILED"; tail -3 gpurun_out/ab.err; continue; }
  tail -1 gpurun_out/ab.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],2), 'value', round(d['value']/1e9,4))"
done 2>&1 | tee gpurun_out/ab_summary.txt
