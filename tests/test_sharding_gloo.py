"""CPU, world_size 2, gloo: the multi-GPU host logic of bench.py (shard by longitude sector,
per-rank work, variable-size gather to rank 0, reassembly in caller order)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    from mops_b200 import sharding, synthetic as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = S.uniform_sphere_seeds(n, 123)
    local, idx = sharding.shard_seeds(seeds, rank, world)
    counts = sharding.all_counts(local.shape[0], world)
    assert sum(counts) == n and counts[rank] == local.shape[0]
    # stand-in for the advection (no GPU here): a deterministic per-particle map
    moved = torch.from_numpy(local * 1.000001 + rank * 0.0)
    parts = sharding.gather_rows(moved, counts, rank, world, dst=0)
    idx_t = torch.from_numpy(idx)
    idx_parts = sharding.gather_rows(idx_t.reshape(-1, 1), counts, rank, world, dst=0)
    if rank == 0:
        full = sharding.scatter_back(n, parts, [p.numpy().reshape(-1) for p in idx_parts])
        np.save(out_path, full)
    else:
        assert parts is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_gather_roundtrip(tmp_path):
    from mops_b200 import sharding, synthetic as S
    n, world = 20000, 2
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    seeds = S.uniform_sphere_seeds(n, 123)
    full = np.load(out)
    assert np.array_equal(full, seeds * 1.000001)
    # sectors are a partition and spatially compact
    sec = sharding.longitude_sector(seeds, 8)
    assert sec.min() == 0 and sec.max() == 7
    cnt = np.bincount(sec, minlength=8)
    assert abs(cnt - n / 8).max() < 0.1 * n / 8


def _snapshot_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    from mops_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fields = _snapshot_fields()
    total = sum(f.size for f in fields)
    chunk, lo, hi = sharding.snapshot_part_bounds(total, rank, world)
    part = torch.from_numpy(sharding.pack_snapshot_part(fields, rank, world))
    assert part.shape[0] == chunk and hi - lo <= chunk
    full = torch.empty(chunk * world, dtype=torch.float64)
    dist.all_gather_into_tensor(full, part)  # NCCL over NVLink in bench.py --snapshot-allgather
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _snapshot_fields():
    rng = np.random.default_rng(5)
    n_cells, L = 1003, 7  # sizes that do not divide evenly
    return [rng.normal(size=(n_cells, L)), rng.normal(size=(n_cells, L)), rng.random((n_cells, L)), rng.random(n_cells)]


def test_partitioned_snapshot_allgather(tmp_path):
    """every rank packs 1/world of the concatenated cell-major fields; the all-gather reproduces the whole snapshot,
    field boundaries and the zero padding of the last chunk included"""
    from mops_b200 import sharding
    world = 3
    out = str(tmp_path / "snap.npy")
    mp.spawn(_snapshot_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    fields = _snapshot_fields()
    flat = np.concatenate([f.reshape(-1) for f in fields])
    full = np.load(out)
    assert np.array_equal(full[:flat.shape[0]], flat)
    assert not full[flat.shape[0]:].any()
    # bounds are a partition of [0, total)
    edges = [sharding.snapshot_part_bounds(flat.shape[0], r, world) for r in range(world)]
    assert edges[0][1] == 0 and edges[-1][2] == flat.shape[0]
    assert all(edges[r][2] == edges[r + 1][1] for r in range(world - 1))


def test_morton_blocks_are_equal_and_contiguous_along_the_curve():
    """product sharding (mops_order_key + mops_shard_bounds; here their Python twins): equal counts for a heavily skewed
    key distribution, every block a contiguous range of the sorted keys, invalid keys (-1) last, every seed exactly once"""
    from mops_b200 import sharding
    rng = np.random.default_rng(4)
    keys = np.concatenate([rng.integers(0, 50, 9000), rng.integers(0, 1_000_000, 1000), np.full(7, -1)]).astype(np.int32)
    rng.shuffle(keys)
    for world in (1, 2, 3, 8):
        blocks = sharding.morton_blocks(keys, world)
        sizes = [b.shape[0] for b in blocks]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == keys.shape[0]
        assert np.array_equal(np.sort(np.concatenate(blocks)), np.arange(keys.shape[0]))
        u = keys.astype(np.uint32)
        for a, b in zip(blocks[:-1], blocks[1:]):
            if a.size and b.size:
                assert u[a].max() <= u[b].min()
        assert (keys[blocks[-1]][-7:] == -1).all()
        for r in range(world):
            lo, hi = sharding.block_bounds(keys.shape[0], r, world)
            assert hi - lo == sizes[r]
