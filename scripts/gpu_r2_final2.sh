# Round 2, final 2-GPU check: full GPU suite (multi-GPU layer on 2 devices) and the driver-shaped N = 2 run with the final defaults
set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu_final_2dev.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_final.log 2> gpurun_out/r02_bench_n2_final.err
tail -1 gpurun_out/r02_bench_n2_final.log | cut -c1-2000; tail -3 gpurun_out/r02_bench_n2_final.err
