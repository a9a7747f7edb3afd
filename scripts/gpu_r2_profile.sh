# Round 2 (1 GPU): the metrics-only ncu pass over one full-size step of the FINAL engine that profiles/k_advect_profile.json
# (bench.py's roofline inputs) is made from, then the driver-shaped bench run.  After the first (segmented) warm-up step the engine
# runs the fresh-seed steps as one launch each: launches 0-2 = warm-up 1, 3 and 4 = warm-ups 2 and 3, 5 = the timed step.
set -x
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
timeout 420 ncu --metrics $M --clock-control none -k regex:k_advect -s 5 -c 1 --csv --log-file gpurun_out/r02_ncu_k_advect_final_metrics.csv $CMD > gpurun_out/r02_ncu_final.log 2>&1
echo "ncu rc=$?"
python scripts/make_profile_json.py gpurun_out/r02_ncu_k_advect_final_metrics.csv gpurun_out/r02_ncu_final.log gpurun_out/k_advect_profile.json | head -30
python bench.py > gpurun_out/r02_bench_n1.log 2> gpurun_out/r02_bench_n1.err; tail -1 gpurun_out/r02_bench_n1.log | cut -c1-700; tail -3 gpurun_out/r02_bench_n1.err
