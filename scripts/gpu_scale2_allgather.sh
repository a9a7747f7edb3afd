# N = 2 (gpurun --gpus 2): the default snapshot distribution against --snapshot-allgather (untested on GPUs when written:
# run this first in the next round; 30-step intervals make the upload the bottleneck, which is what the option removes)
set -x
nvidia-smi -L
for extra in "" "--snapshot-allgather"; do
  for steps in 120 30; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --interval-steps $steps $extra > gpurun_out/bench_n2.log 2>&1 || { echo "FAILED $extra $steps"; tail -5 gpurun_out/bench_n2.log; continue; }
    tail -1 gpurun_out/bench_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$extra] interval $steps:', 'value', round(d['value']/1e9,3), 'ms_per_step', round(d['ms_per_step'],1), 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],1))"
  done
done 2>&1 | tee gpurun_out/allgather_summary.txt
