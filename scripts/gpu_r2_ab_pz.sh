# Round 2, last kernel A/B (1 GPU): zTop packed into the w slot of the velocity records (NOW = 2) and explicit global loads of the
# laundered cell record, against the shipped engine; the L2 access-policy window on / off; parity suite on the winner
set -x
mkdir -p gpurun_out
run() { MOPS_B200_LIB=$PWD/$1 timeout 200 python bench.py --level 8 --particles 16000000 --interval-steps 120 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary 2> gpurun_out/ab.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],2), 'value', round(d['value']/1e9,4))"; }
{ run build_variants/base.so; run build_variants/pz.so; run build_variants/pz_ldg.so; MOPS_NO_L2_WINDOW=1 run build_variants/pz.so no_l2_window; } | tee gpurun_out/r02_ab_pz.txt
BEST=$(python - <<'PY'
best=None
for l in open('gpurun_out/r02_ab_pz.txt'):
    f=l.split()
    if len(f)>=4 and 'no_l2_window' not in l and 'base.so' not in l:
        ms=float(f[f.index('kernel_ms')+1])
        if best is None or ms<best[0]: best=(ms,f[0])
print(best[1] if best else 'build_variants/pz.so')
PY
)
echo "parity suite on $BEST"
( MOPS_B200_LIB=$PWD/$BEST timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5 | tee gpurun_out/r02_pytest_pz.txt
echo "$BEST" > gpurun_out/r02_ab_pz_best.txt
