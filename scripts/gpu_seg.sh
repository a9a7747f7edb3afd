# Compacting multi-launch form (MOPS_SEGMENT_STEPS): parity tests with segments of 7 steps, then timing vs one launch
set -x
MOPS_SEGMENT_STEPS=7 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for s in 0 20 30 40 60; do
  MOPS_SEGMENT_STEPS=$s timeout 200 python bench.py --level 8 --particles 16000000 --interval-steps 120 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/seg.log 2> gpurun_out/seg.err || { echo "seg $s FAILED"; tail -3 gpurun_out/seg.err; continue; }
  tail -1 gpurun_out/seg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('segment_steps $s', 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],2), 'value', round(d['value']/1e9,4), 'launches', d['gpu_launches'])"
done 2>&1 | tee gpurun_out/seg_summary.txt
