# Round 2, call 1 (1 GPU): confirm the committed engine on a fresh box, time the queued build flags, re-capture
# the FINAL advection kernel at full size with a metrics-only ncu pass, and keep a compute-sanitizer log.
#   here (CPU):  bash scripts/build_variants.sh
#   then:        gpurun --timeout 1500 -- 'bash scripts/gpu_r2_call1.sh'
set -x
mkdir -p gpurun_out
./scripts/micro/fp64_lat | tee gpurun_out/r02_fp64_microbench.txt
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu.txt
bash scripts/gpu_ab.sh
cp gpurun_out/ab_summary.txt gpurun_out/r02_ab_flags.txt
for f in build_variants/base.so build_variants/sq_filter.so; do
  [ -f $f ] && MOPS_B200_LIB=$PWD/$f timeout 150 python scripts/bench_secondary.py 2>/dev/null | grep "C3 streamline" | cut -c1-300
done | tee gpurun_out/r02_sq_filter.txt
# metrics-only capture of one full-size SEG launch (no --set full: the 70 GB resident set makes every replay pass expensive)
M=dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,launch__registers_per_thread
timeout 600 ncu --metrics $M --clock-control none -k regex:k_advect -s 9 -c 3 --csv --log-file gpurun_out/r02_ncu_k_advect_fullsize_metrics.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_fullsize.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_ncu_fullsize.log | cut -c1-300
# compute-sanitizer: memcheck, synccheck (partial-warp shuffle masks of k_locate), racecheck
python scripts/sanitize_case.py > gpurun_out/r02_sanitize_plain.log 2>&1 && tail -2 gpurun_out/r02_sanitize_plain.log
for tool in memcheck synccheck racecheck; do
  MOPS_SEGMENT_STEPS=7 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|sanitize case done" gpurun_out/r02_sanitize_$tool.log | head -5
done
ls -la gpurun_out/ | head -50
