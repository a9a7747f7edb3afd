# Round 2, call 3 (1 GPU): full GPU test-suite on the new engine, bench smoke (reduced), source-level ncu capture of the
# hexagon-path pathline kernel on the reduced-footprint launch (level-7 mesh x 80 layers, 4 M particles = 24 per cell, 30 steps)
set -x
mkdir -p gpurun_out
( python -m pytest tests -m gpu -q ) 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_b.txt
CMD="python bench.py --level 7 --particles 4000000 --interval-steps 30 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
$CMD > gpurun_out/plain_small.log 2> gpurun_out/plain_small.err; tail -1 gpurun_out/plain_small.log | cut -c1-300; tail -3 gpurun_out/plain_small.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_advect -s 3 -c 1 -f -o gpurun_out/r02_prof_advect_small $CMD > gpurun_out/ncu_full_small.log 2>&1
tail -2 gpurun_out/ncu_full_small.log | cut -c1-200
python bench.py --level 8 --particles 16000000 --interval-steps 120 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_l8.log 2> gpurun_out/bench_l8.err; tail -1 gpurun_out/bench_l8.log | cut -c1-1500; tail -3 gpurun_out/bench_l8.err
ls -la gpurun_out | head -40
