"""CPU: mops_b200.export writes the same bytes as the reference's tutorial/export_pathline_binary.py
(fixtures in tests/golden/export, made by make_export_golden.py from the unmodified reference script)."""
import filecmp
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = os.path.join(HERE, "golden", "export")


def test_binary_and_json_match_reference_files(tmp_path):
    from make_export_golden import demo_lines
    from mops_b200 import export as X
    lines = demo_lines()
    m0 = X.export_pathlines_to_binary(lines, str(tmp_path / "plain.bin"))
    m1 = X.export_pathlines_to_binary(lines, str(tmp_path / "full.bin"), include_velocity=True, include_scalars=True)
    X.export_pathlines_to_json(lines, str(tmp_path / "all.json"))
    X.export_pathlines_to_json([l for i, l in enumerate(lines) if i not in (2, 5)], str(tmp_path / "dec3.json"), decimation_factor=3)
    for f in ("plain.bin", "plain.meta.json", "full.bin", "full.meta.json", "all.json", "dec3.json"):
        assert filecmp.cmp(str(tmp_path / f), os.path.join(GOLD, f), shallow=False), f
    assert m0["num_particles"] == 7 and m1["fields"] == ["lat", "lon", "velocity_u", "velocity_v", "speed", "temperature", "salinity"]
    assert m0 == json.load(open(os.path.join(GOLD, "plain.meta.json")))


def test_binary_layout_round_trip(tmp_path):
    from mops_b200 import export as X
    rng = np.random.default_rng(5)
    lines = []
    for n in (3, 0, 200):
        d = rng.normal(size=(n, 3))
        lines.append({"points": d / np.linalg.norm(d, axis=1, keepdims=True) * 6370000.0 if n else d, "velocity": rng.normal(size=(n, 3))})
    meta = X.export_pathlines_to_binary(lines, str(tmp_path / "a.bin"), include_velocity=True)
    raw = open(tmp_path / "a.bin", "rb").read()
    assert np.frombuffer(raw[:4], "<i4")[0] == 3
    for line, off in zip(lines, meta["particle_offsets"]):
        n = int(np.frombuffer(raw[off["start"]:off["start"] + 4], "<i4")[0])
        assert n == off["points"] == len(line["points"])
        rows = np.frombuffer(raw[off["start"] + 4: off["start"] + 4 + n * 40], "<f8").reshape(n, 5)
        if n:
            lat, lon, depth = X.xyz_to_lat_lon_depth(*np.asarray(line["points"]).T)
            assert np.array_equal(rows[:, 0], lat) and np.array_equal(rows[:, 1], lon)
            assert np.allclose(depth, 1000.0) and np.allclose(rows[:, 4], np.linalg.norm(line["velocity"], axis=1), rtol=1e-15)
    assert len(raw) == 4 + sum(4 + o["points"] * 40 for o in meta["particle_offsets"])
