set -x
nvidia-smi -L | wc -l; free -g | head -2; nproc
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/bench_n8.log 2>&1
tail -c 3000 gpurun_out/bench_n8.log
