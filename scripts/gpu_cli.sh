set -x
timeout 600 python -m pytest tests/test_cpp_api_gpu.py -m gpu -x -q -s -k "cli or mpas_files" > gpurun_out/cli.log 2>&1
grep -E "passed|failed|Error|assert|== " gpurun_out/cli.log | tail -20 | cut -c1-200
