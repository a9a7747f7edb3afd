# Round 2 (1 GPU): the driver-shaped bench run of the final engine, then the metrics-only ncu pass over one full-size step that
# profiles/k_advect_profile.json (bench.py's roofline inputs) is made from, then the ncu launch list of the same command
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.log 2> gpurun_out/r02_bench_n1.err; tail -1 gpurun_out/r02_bench_n1.log | cut -c1-3000; tail -3 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.log 2>&1; tail -1 gpurun_out/r02_bench_ref.log | cut -c1-600
M=dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
timeout 900 ncu --metrics $M --clock-control none -k regex:k_advect -s 9 -c 3 --csv --log-file gpurun_out/r02_ncu_k_advect_final_metrics.csv $CMD > gpurun_out/r02_ncu_final.log 2>&1
echo "ncu rc=$?"
python scripts/make_profile_json.py gpurun_out/r02_ncu_k_advect_final_metrics.csv gpurun_out/r02_ncu_final.log gpurun_out/k_advect_profile.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu list rc=$?"
python scripts/bench_secondary.py > gpurun_out/r02_secondary_c2_c3.jsonl 2> gpurun_out/secondary.err; cut -c1-400 gpurun_out/r02_secondary_c2_c3.jsonl
