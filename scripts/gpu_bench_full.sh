set -x
free -g | head -2; nproc
( time python bench.py ) > gpurun_out/bench_full.log 2>&1
tail -c 6000 gpurun_out/bench_full.log
