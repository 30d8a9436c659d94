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
    # shard_of_replica is the inverse of shard_bounds for every split, divisible or not, R < G included
    for Rr in range(1, 40):
        for Gg in range(1, 12):
            for rank in range(Gg):
                lo, hi = parallel.shard_bounds(Rr, rank, Gg)
                assert all(parallel.shard_of_replica(Rr, Gg, r) == rank for r in range(lo, hi)), (Rr, Gg, rank)
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
    # counts as int64 (columns 0 and 3), sums as float64: a count beyond 2**53 stays exact
    mixed = torch.tensor([[2.0 ** 53, 0.25, 0.5, 7.0]], dtype=torch.float64) if rank == 0 else \
        torch.tensor([[1.0, 0.25, 0.5, 5.0]], dtype=torch.float64)
    counts = mixed[:, [0, 3]].to(torch.int64).clone()
    dist.all_reduce(counts)
    parallel.all_reduce_tallies(mixed, count_cols=[0, 3])
    if rank == 0:
        np.save(os.path.join(tmp, "merged.npy"), merged.numpy())
        np.save(os.path.join(tmp, "shared.npy"), shared.numpy())
        np.save(os.path.join(tmp, "mixed.npy"), mixed.numpy())
        np.save(os.path.join(tmp, "counts.npy"), counts.numpy())
    dist.destroy_process_group()


def test_tally_all_reduce_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    merged = np.load(tmp_path / "merged.npy")
    base = np.arange(12, dtype=float).reshape(3, 4)
    assert np.array_equal(merged, np.vstack([base, base + 100]))
    assert np.array_equal(np.load(tmp_path / "shared.npy"), np.full((2, 4), 3.0))
    mixed, counts = np.load(tmp_path / "mixed.npy"), np.load(tmp_path / "counts.npy")
    assert counts.tolist() == [[2 ** 53 + 1, 12]]          # the int64 path is exact ...
    assert mixed[0, 1] == 0.5 and mixed[0, 2] == 1.0 and mixed[0, 3] == 12.0
    assert parallel.count_columns(3, 2) == [0, 1, 2, 9, 10]
