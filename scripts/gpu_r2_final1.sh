# Round 2, final single-GPU check: the driver's sequence (GPU tests, smoke, default bench) on the committed engine, the
# one-launch vs compacting A/B, and C5's 720-step interval at N = 1 (device arm)
set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1_final.log 2> gpurun_out/r02_bench_n1_final.err; tail -1 gpurun_out/r02_bench_n1_final.log | cut -c1-900; tail -3 gpurun_out/r02_bench_n1_final.err
for seg in 0 40; do
  MOPS_SEGMENT_STEPS=$seg python bench.py --level 8 --particles 16000000 --interval-steps 120 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('segment_steps=$seg kernel_ms', round(d['roofline']['kernel_ms_per_launch'],2), 'value', round(d['value']/1e9,4))"
done | tee gpurun_out/r02_ab_segment.txt
python bench.py --interval-steps 720 --steps 2 --warmup 1 --no-e2e --no-secondary --no-cpu-baseline > gpurun_out/r02_c5_720_n1.log 2> gpurun_out/r02_c5_720_n1.err; tail -1 gpurun_out/r02_c5_720_n1.log | cut -c1-900; tail -2 gpurun_out/r02_c5_720_n1.err
