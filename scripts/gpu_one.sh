set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "$1" > gpurun_out/one.log 2>&1
tail -25 gpurun_out/one.log | cut -c1-250
