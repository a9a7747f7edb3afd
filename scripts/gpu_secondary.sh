set -x
python -m pytest tests -m gpu -x -q -s -k "large_mesh" 2>&1 | grep -E "large|passed|failed|Error" | tail
python scripts/bench_secondary.py > gpurun_out/secondary.jsonl 2> gpurun_out/secondary.err
cat gpurun_out/secondary.jsonl; tail -3 gpurun_out/secondary.err
