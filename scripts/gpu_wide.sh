set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "voronoi or pinned" > gpurun_out/wide.log 2>&1
grep -E "passed|failed|Error|assert|\[m8|\[m20" gpurun_out/wide.log | tail -20 | cut -c1-200
