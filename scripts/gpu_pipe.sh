set -x
timeout 900 python -m pytest tests/test_edge_cases_gpu.py -m gpu -x -q > gpurun_out/edge.log 2>&1
tail -15 gpurun_out/edge.log | cut -c1-220
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep-host-chunk 0,2097152,4194304,8388608,16777216 > gpurun_out/pipe_bench.log 2> gpurun_out/pipe_bench.err
tail -3 gpurun_out/pipe_bench.log | cut -c1-1500
grep sweep gpurun_out/pipe_bench.err
tail -5 gpurun_out/pipe_bench.err | cut -c1-300
