"""Trajectory export for web viewers: the BIN (+ .meta.json) and JSON formats of the reference's
`tutorial/export_pathline_binary.py:26-215`, written with whole-array numpy operations instead of one
`struct.pack` per value.  Byte-for-byte equal to the reference script's files (tests/test_export.py, against
fixtures made by running that script, tests/golden/make_export_golden.py).

A line is the dict pyMOPS returns: {"points": [n][3] ECEF metres, "velocity": [n][3], "temperature": [n],
"salinity": [n]}.  Entries that are not dicts are skipped the way the reference skips them: counted in the
header, no record, no offset entry.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Sequence

import numpy as np

EARTH_RADIUS_EXPORT = 6_371_000.0  # export_pathline_binary.py:17 (not the 6,371,010 of the sampling code)


def xyz_to_lat_lon_depth(x, y, z, R: float = EARTH_RADIUS_EXPORT):
    """ECEF -> (lat deg, lon deg, depth m); export_pathline_binary.py:17-23"""
    lon = np.degrees(np.arctan2(y, x))
    r = np.sqrt(x * x + y * y + z * z)
    lat = np.degrees(np.arcsin(z / r))
    return lat, lon, R - r


def _points(line) -> np.ndarray:
    return np.asarray(line.get("points", []))


def export_pathlines_to_binary(trajectory_lines: Sequence[Any], output_path: str, include_velocity: bool = False,
                               include_scalars: bool = False) -> Dict[str, Any]:
    """little-endian: int32 line count, then per line int32 n followed by n rows of float64
    (lat, lon[, vel_x, vel_y, |vel|][, temperature, salinity]); metadata with byte offsets goes to
    <output>.meta.json (export_pathline_binary.py:26-131)."""
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    fields = ["lat", "lon"]
    if include_velocity:
        fields += ["velocity_u", "velocity_v", "speed"]
    if include_scalars:
        fields += ["temperature", "salinity"]
    offsets: List[Dict[str, int]] = []
    chunks: List[bytes] = [np.int32(len(trajectory_lines)).astype("<i4").tobytes()]
    at = 4
    for line in trajectory_lines:
        if not isinstance(line, dict):
            continue
        P = _points(line)
        n = P.shape[0]
        if n < 1:
            chunks.append(np.zeros(1, "<i4").tobytes())
            offsets.append({"start": at, "points": 0})
            at += 4
            continue
        lat, lon, _ = xyz_to_lat_lon_depth(P[:, 0], P[:, 1], P[:, 2])
        cols = [lat, lon]
        if include_velocity:
            V = np.asarray(line.get("velocity", []))
            if V.shape[0] == n:
                V = V.reshape(n, -1).astype(np.float64)
                cols.append(V[:, 0] if V.shape[1] > 0 else np.zeros(n))
                cols.append(V[:, 1] if V.shape[1] > 1 else np.zeros(n))
                cols.append(np.array([np.linalg.norm(V[i]) for i in range(n)]))  # per row: BLAS dot rounding, as the reference
            else:
                cols += [np.zeros(n)] * 3
        if include_scalars:
            for key in ("temperature", "salinity"):
                a = np.asarray(line.get(key, []), dtype=np.float64).reshape(-1)
                c = np.zeros(n)
                m = min(n, a.shape[0])
                c[:m] = a[:m]
                cols.append(c)
        rows = np.ascontiguousarray(np.column_stack(cols).astype("<f8"))
        chunks.append(np.int32(n).astype("<i4").tobytes())
        chunks.append(rows.tobytes())
        offsets.append({"start": at, "points": int(n)})
        at += 4 + rows.nbytes
    with open(output_path, "wb") as f:
        for c in chunks:
            f.write(c)
    metadata = {"format_version": "1.0", "num_particles": len(trajectory_lines), "fields": fields, "data_type": "float64",
                "byte_order": "little", "particle_offsets": offsets}
    with open(output_path.with_suffix(".meta.json"), "w") as f:
        json.dump(metadata, f, indent=2)
    return metadata


def export_pathlines_to_json(trajectory_lines: Sequence[Any], output_path: str, decimation_factor: int = 1) -> None:
    """{"format": "mops-pathlines-v1", "particles": [{"id", "points": [[lat, lon]..], "velocity", "temperature",
    "salinity"}]}; lines without points are dropped, every decimation_factor-th point is kept
    (export_pathline_binary.py:134-215)."""
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    particles = []
    for idx, line in enumerate(trajectory_lines):
        if not isinstance(line, dict):
            continue
        P = _points(line)
        if P.shape[0] < 1:
            continue
        lat, lon, _ = xyz_to_lat_lon_depth(P[:, 0], P[:, 1], P[:, 2])
        keep = None
        if decimation_factor > 1:
            keep = np.arange(0, len(lat), decimation_factor)
            lat, lon = lat[keep], lon[keep]
        n = len(lat)
        particle: Dict[str, Any] = {"id": idx, "points": np.column_stack([lat, lon]).astype(float).tolist()}
        V = line.get("velocity")
        if V is not None and len(V) > 0:
            V = np.asarray(V)
            if keep is not None:
                V = V[keep]
            particle["velocity"] = [[float(v) for v in V[i]] for i in range(min(len(V), n))]
        for key in ("temperature", "salinity"):
            a = line.get(key)
            if a is not None and len(a) > 0:
                a = np.asarray(a)
                if keep is not None:
                    a = a[keep]
                particle[key] = [float(t) for t in a[:n]]
        particles.append(particle)
    with open(output_path, "w") as f:
        json.dump({"format": "mops-pathlines-v1", "num_particles": len(particles), "particles": particles}, f, indent=2)
