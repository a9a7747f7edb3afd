#!/usr/bin/env python
"""profiles/k_advect_profile.json from a metrics-only ncu pass over the k_advect launches of ONE full-size bench step
(scripts/gpu_r2_profile.sh): fp64-pipe thread instructions per executed particle-step and DRAM bytes per launch.
bench.py reads the file for its roofline block (fp64-pipe fraction = instructions/step x measured steps/s / pipe peak)."""
import csv
import json
import sys

csv_path, bench_log, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(csv_path)))
hdr = None
per = {}
name = None
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if "k_advect" not in d["Kernel Name"]:
            continue
        name = d["Kernel Name"]
        per.setdefault(d["ID"], {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
line = json.loads([l for l in open(bench_log) if l.startswith("{")][-1])
cfg = line["config"]
steps_per_bench_step = cfg["executed_fraction"] * cfg["particles_total"] * cfg["interval_steps"]
launches = sorted(per, key=int)
tot = lambda m: sum(per[i].get(m, 0.0) for i in launches)
fp64 = tot("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum") + tot("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum") + \
       tot("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
out = {
    "kernel": name,
    "source": f"ncu metrics-only pass over the {len(launches)} k_advect launches of one full-size bench step "
              f"({cfg['particles_total']} seeds x {cfg['interval_steps']} steps, {cfg['cells']} cells x {cfg['layers']} layers); "
              f"{csv_path.split('/')[-1]}",
    "launches_per_step": len(launches),
    "particle_steps_in_capture": steps_per_bench_step,
    "fp64_thread_inst_per_particle_step": fp64 / steps_per_bench_step,
    "fp64_note": "DFMA + DMUL + DADD thread instructions (DSETP, ~3 % more pipe slots, not counted)",
    "thread_inst_per_particle_step": tot("smsp__thread_inst_executed.sum") / steps_per_bench_step,
    "dram_bytes_per_launch": (tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")) / len(launches),
    "dram_bytes_per_particle_step": (tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")) / steps_per_bench_step,
    "fp64_pipe_active_pct_ncu": sum(per[i].get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for i in launches) / len(launches),
    "l1_hit_pct": sum(per[i].get("l1tex__t_sector_hit_rate.pct", 0.0) for i in launches) / len(launches),
    "l2_hit_pct": sum(per[i].get("lts__t_sector_hit_rate.pct", 0.0) for i in launches) / len(launches),
    "l1_data_pipe_pct": sum(per[i].get("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", 0.0) for i in launches) / len(launches),
    "registers_per_thread": per[launches[0]].get("launch__registers_per_thread"),
    "warps_active_pct": sum(per[i].get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) for i in launches) / len(launches),
    "kernel_ms_under_ncu": [per[i].get("gpu__time_duration.sum", 0.0) / 1e6 for i in launches],
}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
