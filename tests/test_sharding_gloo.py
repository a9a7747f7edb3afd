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
