"""Golden files for mops_b200/export.py: runs the reference's own tutorial/export_pathline_binary.py (imported
from /root/reference, this container only) on deterministic lines and stores its output files.

    python tests/golden/make_export_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def demo_lines():
    """7 entries: full lines, a line without velocity, an empty line, a non-dict, a short scalar list"""
    rng = np.random.default_rng(20261018)
    out = []
    for k, n in enumerate((5, 1, 12, 0, 7)):
        d = rng.normal(size=(n, 3))
        p = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-300) * (6371010.0 - 100.0 * k) if n else np.zeros((0, 3))
        out.append({"points": p.tolist(), "velocity": rng.normal(size=(n, 3)).tolist(), "temperature": rng.uniform(0, 30, n).tolist(),
                    "salinity": rng.uniform(30, 38, n).tolist()})
    out[2].pop("velocity")
    out[4]["temperature"] = out[4]["temperature"][:3]
    out.insert(3, None)
    out.append({"points": np.array([[6371010.0, 0.0, 0.0], [0.0, 0.0, -6371010.0]]), "velocity": np.zeros((2, 3))})
    return out


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference/tutorial")
    import export_pathline_binary as ref  # the reference script, unmodified
    out = os.path.join(HERE, "export")
    os.makedirs(out, exist_ok=True)
    lines = demo_lines()
    ref.export_pathlines_to_binary(lines, os.path.join(out, "plain.bin"))
    ref.export_pathlines_to_binary(lines, os.path.join(out, "full.bin"), include_velocity=True, include_scalars=True)
    ref.export_pathlines_to_json(lines, os.path.join(out, "all.json"))
    ref.export_pathlines_to_json([l for i, l in enumerate(lines) if i not in (2, 5)], os.path.join(out, "dec3.json"), decimation_factor=3)
