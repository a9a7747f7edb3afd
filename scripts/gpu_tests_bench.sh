set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
tail -3 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_try.log 2> gpurun_out/bench_try.err
tail -1 gpurun_out/bench_try.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch'], d['clocks'])"
