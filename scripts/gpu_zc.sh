set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/zc.log 2>&1; tail -3 gpurun_out/zc.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value %.4g ms/step %.1f | e2e %.4g ms/step %.1f | kernel_ms %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_launch']))"
