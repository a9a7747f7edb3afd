"""GPU: the single-process multi-GPU layer (mops_multi_*): seeds sorted along the mesh's Morton curve on device 0, equal
contiguous blocks shipped to the devices over NCCL, blocks integrated concurrently, records / end points / status gathered
back in caller order over NCCL.  Results must not depend on the number of devices: every output is compared bit for bit
with the single-device call and, through it, with the oracle.  With one visible GPU the same code runs as a world of one
(no NCCL traffic); the N > 1 cases need >= 2 GPUs (gpurun --gpus 2) and are skipped otherwise."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _case():
    m = cases.mesh(4)
    s0, s1 = cases.snapshots(4, 12, "rich")
    seeds = cases.seeds_random(6000, seed=77)
    seeds[11] = np.nan
    depths = np.linspace(5.0, 900.0, seeds.shape[0]).astype(np.float32)
    return m, s0, s1, seeds, depths


@pytest.mark.parametrize("n_dev", [1, 2, 0])
def test_multi_matches_single_device(n_dev):
    from mops_b200 import capi
    from oracle import port_oracle as P
    have = _ngpu()
    if n_dev == 2 and have < 2:
        pytest.skip("needs 2 GPUs")
    if n_dev == 0 and have < 2:
        pytest.skip("all-devices case needs >= 2 GPUs")
    m, s0, s1, seeds, depths = _case()
    e = capi.Engine(0)
    e.set_mesh(m); e.set_snapshot(0, s0); e.set_snapshot(1, s1)
    cells = P.locate(m, seeds)
    cells_bad = cells.copy(); cells_bad[40::97] = -1
    mm = capi.MultiEngine(n_dev)
    try:
        assert mm.n == (have if n_dev == 0 else n_dev)
        mm.set_mesh(m); mm.set_snapshot(0, s0); mm.set_snapshot(1, s1, async_=True)
        for cell0 in (None, cells_bad):
            a = e.pathline(0, 1, seeds, 300, 21600, 3600, depths=depths, cell0=cell0, want_attr=True)
            b = mm.pathline(0, 1, seeds, 300, 21600, 3600, depths=depths, cell0=cell0, want_attr=True)
            for k in ("raw_pos", "raw_vel", "raw_attr", "pos", "depth", "status", "steps_alive", "final_cell"):
                assert np.array_equal(a[k], b[k], equal_nan=True), (n_dev, k)
            for f in ("particle_steps", "alive_at_end"):
                assert int(getattr(a["stats"], f)) == int(getattr(b["stats"], f)), f
        a = e.streamline(0, seeds, 300, 21600, 1800, depth=300.0, method="euler")
        b = mm.streamline(0, seeds, 300, 21600, 1800, depth=300.0, method="euler")
        for k in ("raw_pos", "raw_vel", "pos", "depth", "status", "steps_alive", "final_cell"):
            assert np.array_equal(a[k], b[k], equal_nan=True), (n_dev, k)
        # and against the oracle directly
        want = P.pathline(m, P.prepare(m, s0), P.prepare(m, s1), seeds, cells, 300, 21600, 3600, depths=depths)
        got = mm.pathline(0, 1, seeds, 300, 21600, 3600, depths=depths, cell0=cells, want_attr=True)
        assert np.array_equal(got["raw_pos"], want["raw_pos"], equal_nan=True) and np.array_equal(got["status"], want["status"])
        # device 0's context serves the single-device calls
        assert np.array_equal(mm.dev0.locate(seeds[:500]), cells[:500])
        # degenerate sizes: fewer particles than devices, empty call
        few = mm.pathline(0, 1, seeds[:1], 300, 21600, 3600, depth=300.0)
        one = e.pathline(0, 1, seeds[:1], 300, 21600, 3600, depth=300.0, want_attr=False)
        assert np.array_equal(few["raw_pos"], one["raw_pos"])
        none = mm.pathline(0, 1, seeds[:0], 300, 21600, 3600, depth=300.0)
        assert none["raw_pos"].shape[0] == 0
    finally:
        mm.close()
        e.close()


def test_cpp_api_on_all_devices(tmp_path):
    """the C++ drop-in (tutorial/pathLine) with MOPS_DEVICES=all gives byte-identical line files to the single-device run"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tutorial", "bin", "pathLine")
    if not os.path.exists(exe):
        pytest.skip("tutorial binaries not built")
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    from mops_b200 import synthetic as S
    m = cases.mesh(5)
    snaps = [S.solid_body_snapshot(m, 20, 0.3 + 0.1 * i, tilt=0.3 + 0.02 * i, shear=0.2, w_amp=1e-3, with_attrs=True) for i in range(3)]
    fx = str(tmp_path / "fx.bin")
    S.dump_fixture(fx, m, snaps)
    outs = {}
    for tag, env in (("one", {}), ("all", {"MOPS_DEVICES": "all"})):
        pre = str(tmp_path / f"lines_{tag}")
        r = subprocess.run([exe, fx, pre], env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = [open(f"{pre}_{s}.bin", "rb").read() for s in range(2)]
        if tag == "all":
            assert "Devices in use" in r.stdout
    assert outs["one"] == outs["all"]
