# Round 2 (8 GPUs): the driver-shaped N = 8 run (default flags), the snapshot-distribution A/B, C5 as written (720-step
# intervals) and C4 (10 M seeds, 29 chained 720-step intervals), plus the C++ tutorial on all devices
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 500 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8.log 2> gpurun_out/r02_bench_n8.err
tail -1 gpurun_out/r02_bench_n8.log | cut -c1-2600; tail -3 gpurun_out/r02_bench_n8.err
timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 2 --snapshot-upload replicated --no-e2e --no-secondary > gpurun_out/r02_bench_n8_repl.log 2> gpurun_out/r02_bench_n8_repl.err
tail -1 gpurun_out/r02_bench_n8_repl.log | cut -c1-400; tail -3 gpurun_out/r02_bench_n8_repl.err
timeout 400 $TR bench.py --gpus 8 --interval-steps 720 --steps 3 --warmup 1 --no-secondary > gpurun_out/r02_c5_720_n8.log 2> gpurun_out/r02_c5_720_n8.err
tail -1 gpurun_out/r02_c5_720_n8.log | cut -c1-2600; tail -3 gpurun_out/r02_c5_720_n8.err
timeout 400 $TR bench.py --gpus 8 --particles 10000000 --interval-steps 720 --steps 29 --warmup 0 --chain --no-secondary > gpurun_out/r02_c4_n8.log 2> gpurun_out/r02_c4_n8.err
tail -1 gpurun_out/r02_c4_n8.log | cut -c1-2600; tail -3 gpurun_out/r02_c4_n8.err
# the C++ drop-in on all 8 devices (tutorial/pathLine, MOPS_DEVICES=all) against the single-device run
python - <<'PY' > gpurun_out/r02_tutorial_8dev.txt 2>&1
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from mops_b200 import synthetic as S
m = S.icosahedral_mesh(6)
snaps = [S.solid_body_snapshot(m, 20, 0.3 + 0.1 * i, tilt=0.3 + 0.02 * i, shear=0.2, w_amp=1e-3, with_attrs=True) for i in range(4)]
S.dump_fixture("/tmp/fx8.bin", m, snaps)
outs = {}
for tag, env in (("one", {}), ("all", {"MOPS_DEVICES": "all"})):
    t = time.time()
    r = subprocess.run(["tutorial/bin/pathLine", "/tmp/fx8.bin", f"/tmp/lines8_{tag}"], env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    print(tag, "rc", r.returncode, "%.1fs" % (time.time() - t)); print(r.stdout[-1500:]); print(r.stderr[-500:])
    outs[tag] = [open(f"/tmp/lines8_{tag}_{s}.bin", "rb").read() for s in range(3)]
print("byte-identical line files, 1 device vs all devices:", outs["one"] == outs["all"])
PY
tail -25 gpurun_out/r02_tutorial_8dev.txt
