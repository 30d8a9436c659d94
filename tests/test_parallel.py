"""N > 1 host logic on CPU: replica sharding and the tally all-reduce over a 2-rank gloo group."""
import os
import sys

import numpy as np
import pytest

from kmos_b200 import parallel


def test_shard_bounds_cover_and_keep_points_together():
    R, G = 16384, 8
    covered = []
    for rank in range(G):
        lo, hi = parallel.shard_bounds(R, rank, G)
        covered += list(range(lo, hi))
        assert (hi - lo) % 64 == 0  # 64 seeds of a sweep point never straddle two GPUs
        assert all(parallel.shard_of_replica(R, G, r) == rank for r in (lo, hi - 1))
    assert covered == list(range(R))
    # ragged split
    sizes = [parallel.shard_bounds(10, r, 4) for r in range(4)]
    assert sizes == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert len(set(parallel.global_seeds(1000).tolist())) == 1000


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank owns 3 local sweep points of 4 words; global table has 6 points
    local = torch.arange(12, dtype=torch.float64).reshape(3, 4) + 100 * rank
    merged = parallel.merge_group_tallies(local, 3 * rank, 6)
    parallel.all_reduce_tallies(merged)
    # a second tally where both ranks contribute to the same groups (seeds split across GPUs)
    shared = torch.full((2, 4), float(rank + 1), dtype=torch.float64)
    parallel.all_reduce_tallies(shared)
    if rank == 0:
        np.save(os.path.join(tmp, "merged.npy"), merged.numpy())
        np.save(os.path.join(tmp, "shared.npy"), shared.numpy())
    dist.destroy_process_group()


def test_tally_all_reduce_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    merged = np.load(tmp_path / "merged.npy")
    base = np.arange(12, dtype=float).reshape(3, 4)
    assert np.array_equal(merged, np.vstack([base, base + 100]))
    assert np.array_equal(np.load(tmp_path / "shared.npy"), np.full((2, 4), 3.0))
