# Round-end confirmation of the committed engine: parity tests, the default bench, a reduced-footprint ncu capture
# of the advection kernel (same command first without ncu), then the secondary C2/C3 measurements.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
( time python bench.py ) > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-300
CMD="python bench.py --level 7 --particles 4000000 --interval-steps 30 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_small.log 2> gpurun_out/plain_small.err && tail -1 gpurun_out/plain_small.log | cut -c1-200 &&
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_advect -s 3 -c 1 -f -o gpurun_out/prof_advect_small $CMD > gpurun_out/ncu_full_small.log 2>&1
tail -2 gpurun_out/ncu_full_small.log | cut -c1-200
timeout 150 python scripts/bench_secondary.py > gpurun_out/secondary.jsonl 2> gpurun_out/secondary.err
cat gpurun_out/secondary.jsonl | cut -c1-400; tail -3 gpurun_out/secondary.err
