"""Particle sharding for the multi-GPU path: particles are independent, the mesh
and the resident snapshots are replicated on every GPU, each rank owns an equal, spatially compact
block of the seed set sorted along the mesh's Morton curve, and the only exchange is the gather of
results to one owner at the end of an interval (in the product: mops_dist_gather_traj / mops_multi_*,
NCCL inside the library).  Host-side logic only; torch.distributed does the bench's plumbing (NCCL on
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def longitude_sector(xyz: np.ndarray, world: int) -> np.ndarray:
    """sector id in [0, world) of every particle: equal-width longitude wedges (round 1's sharding; balanced only for
    uniform seeds -- bench.py and the product use morton_blocks / mops_order_key now; kept for engine-less callers)"""
    lon = np.arctan2(xyz[:, 1], xyz[:, 0])
    return np.minimum(((lon + np.pi) / (2.0 * np.pi) * world).astype(np.int64), world - 1)


def shard_seeds(xyz: np.ndarray, rank: int, world: int) -> Tuple[np.ndarray, np.ndarray]:
    """(local seeds, their indices in the global seed array); world == 1 keeps everything"""
    if world <= 1:
        return xyz, np.arange(xyz.shape[0], dtype=np.int64)
    idx = np.nonzero(longitude_sector(xyz, world) == rank)[0]
    return np.ascontiguousarray(xyz[idx]), idx


def block_bounds(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's block when n_total key-sorted seeds are cut into `world` equal contiguous blocks (the first
    n_total % world blocks hold one more) -- the Python twin of mops_shard_bounds (include/mops_b200.h)"""
    world = max(world, 1)
    base, extra = divmod(int(n_total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def morton_blocks(keys: np.ndarray, world: int) -> List[np.ndarray]:
    """caller indices of every rank's block: a stable sort of the seed set by processing-order key (the rank of each
    seed's cell along the mesh's Morton curve, mops_order_key), cut by block_bounds.  Equal counts for ANY seed
    distribution; each block is a spatially compact range of the curve."""
    order = np.argsort(np.asarray(keys).astype(np.uint32), kind="stable")  # -1 (no cell) sorts last, as on the device
    return [order[slice(*block_bounds(order.shape[0], r, world))] for r in range(max(world, 1))]


def all_counts(n_local: int, world: int, device=None) -> List[int]:
    import torch
    import torch.distributed as dist
    if world <= 1:
        return [n_local]
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    out = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, t)
    return [int(c.item()) for c in out]


def gather_rows(local, counts: List[int], rank: int, world: int, dst: int = 0):
    """gather a [n_local, k] tensor of every rank to `dst`; n_local may differ between ranks
    (gather collectives need equal shapes, so shards are padded to the largest and sliced on
    arrival).  Returns the list of per-rank tensors on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    if world <= 1:
        return [local]
    cmax = max(counts)
    tail = tuple(local.shape[1:])
    if local.shape[0] == cmax:
        padded = local.contiguous()
    else:
        padded = torch.empty((cmax,) + tail, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]].copy_(local)
    if rank == dst:
        bufs = [torch.empty((cmax,) + tail, dtype=local.dtype, device=local.device) for _ in counts]
        dist.gather(padded, bufs, dst=dst)
        return [b[:c] for b, c in zip(bufs, counts)]
    dist.gather(padded, None, dst=dst)
    return None


def snapshot_part_bounds(total: int, rank: int, world: int) -> Tuple[int, int, int]:
    """The next snapshot does not have to cross every rank's PCIe link whole: its cell-major fields, concatenated
    (zonal, meridional, layerThickness, bottomDepth = `total` doubles), are cut into `world` chunks of `chunk` doubles
    (the last one zero-padded); rank r uploads [lo, hi) and an all-gather over NVLink completes the copy on every GPU.
    Returns (chunk, lo, hi)."""
    chunk = -(-total // max(world, 1))
    lo = min(rank * chunk, total)
    hi = min(lo + chunk, total)
    return chunk, lo, hi


def pack_snapshot_part(fields, rank: int, world: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    """this rank's chunk of the concatenation of `fields` (arrays, flattened in C order), zero-padded to the chunk
    length, without materialising the concatenation"""
    flats = [np.asarray(f, dtype=np.float64).reshape(-1) for f in fields]
    total = int(sum(f.shape[0] for f in flats))
    chunk, lo, hi = snapshot_part_bounds(total, rank, world)
    if out is None:
        out = np.empty(chunk, dtype=np.float64)
    assert out.shape == (chunk,)
    out[hi - lo:] = 0.0
    start = 0
    for f in flats:
        a, b = max(lo, start), min(hi, start + f.shape[0])
        if a < b:
            out[a - lo:b - lo] = f[a - start:b - start]
        start += f.shape[0]
    return out


def scatter_back(global_n: int, parts, indices: List[np.ndarray]) -> np.ndarray:
    """reassemble gathered per-rank rows into caller (global seed) order"""
    import torch
    first = parts[0]
    out = np.empty((global_n,) + tuple(first.shape[1:]), dtype=first.cpu().numpy().dtype)
    for p, idx in zip(parts, indices):
        out[idx] = p.cpu().numpy()
    return out
