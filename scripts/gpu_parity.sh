set -x
nvidia-smi -L
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s 2>&1 | tail -60
