set -x
timeout 900 python -m pytest tests/test_edge_cases_gpu.py -m gpu -x -q > gpurun_out/edge.log 2>&1
tail -30 gpurun_out/edge.log | cut -c1-220
