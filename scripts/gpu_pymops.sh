set -x
MOPS_BACKTRACE=1 timeout 300 python -X faulthandler -m pytest tests/test_pymops_gpu.py -m gpu -x -q -s > gpurun_out/pymops.log 2>&1
grep -v "^  File\|^$" gpurun_out/pymops.log | tail -40 | cut -c1-220
