set -x
python -m pytest tests -m gpu -x -q -s > gpurun_out/gpu_tests.log 2>&1
grep -E "interval|C1:|passed|failed|bit-identical" gpurun_out/gpu_tests.log | tail -40
