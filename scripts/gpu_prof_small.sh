set -x
CMD="python bench.py --level 8 --particles 16000000 --interval-steps 30 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/small_plain.log 2>&1 || { tail -5 gpurun_out/small_plain.log; exit 1; }
tail -1 gpurun_out/small_plain.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_advect -s 3 -c 1 -f -o gpurun_out/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
tail -3 gpurun_out/ncu_small.log | cut -c1-200
ls -la gpurun_out/prof_small.ncu-rep
