for b in 2 3; do
  echo "== MINB=$b"
  MOPS_ADVECT_MINB=$b python bench.py --steps 3 --warmup 3 --interval-steps 30 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value %.4g  ms/step %.1f kernel_ms %.1f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch']))"
done
echo "== unsorted"
python bench.py --steps 3 --warmup 3 --interval-steps 30 --no-cpu-baseline --no-e2e --no-sort 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value %.4g  ms/step %.1f kernel_ms %.1f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch']))"
