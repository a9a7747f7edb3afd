# Round 2, final N-GPU run with the final defaults (usage: gpurun --gpus N -- 'bash scripts/gpu_r2_final8.sh N')
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 500 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n${N}_final.log 2> gpurun_out/r02_bench_n${N}_final.err
tail -1 gpurun_out/r02_bench_n${N}_final.log | cut -c1-2200; tail -3 gpurun_out/r02_bench_n${N}_final.err
