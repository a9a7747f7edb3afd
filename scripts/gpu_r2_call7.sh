# Round 2, call 7 (2 GPUs): full GPU test-suite (incl. the multi-GPU layer on 2 devices and the benchmark-shape parity tests),
# bench at N = 2 under torchrun (Morton-block sharding, in-library NCCL gather of trajectories, 1/N snapshot upload + all-gather)
set -x
mkdir -p gpurun_out
nvidia-smi -L
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12 | tee gpurun_out/r02_pytest_gpu_2dev.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --level 8 --particles 16000000 --interval-steps 120 --steps 3 --warmup 2 > gpurun_out/r02_bench_n2_l8.log 2> gpurun_out/r02_bench_n2_l8.err
tail -1 gpurun_out/r02_bench_n2_l8.log | cut -c1-1200; tail -5 gpurun_out/r02_bench_n2_l8.err
timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 2 > gpurun_out/r02_bench_n2.log 2> gpurun_out/r02_bench_n2.err
tail -1 gpurun_out/r02_bench_n2.log | cut -c1-2500; tail -5 gpurun_out/r02_bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 2 --snapshot-upload replicated --no-e2e > gpurun_out/r02_bench_n2_repl.log 2> gpurun_out/r02_bench_n2_repl.err
tail -1 gpurun_out/r02_bench_n2_repl.log | cut -c1-600; tail -3 gpurun_out/r02_bench_n2_repl.err
