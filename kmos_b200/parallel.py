"""Multi-GPU plumbing: replicas are sharded over ranks (one process per GPU), stepping needs no traffic, and
one collective per sampling point combines the tallies.

The reference's analogue is ``ModelRunner`` (kmos/run/__init__.py:2005-2366): a process pool over parameter
points whose workers append rows to a shared ``.dat`` file under a lock file.
"""
import numpy as np


def shard_bounds(n_replicas, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: replica r lives on GPU floor(r*G/R), so the seeds of one
    sweep point stay together whenever seeds_per_point divides R/G."""
    lo = (n_replicas * rank) // world
    hi = (n_replicas * (rank + 1)) // world
    return lo, hi


def shard_of_replica(n_replicas, world, r):
    """Rank whose shard_bounds block holds replica r: the smallest k with (n_replicas*(k+1))//world > r."""
    k = (r * world) // n_replicas if n_replicas > 0 else 0
    while k + 1 < world and (n_replicas * (k + 1)) // world <= r:
        k += 1
    while k > 0 and (n_replicas * k) // world > r:
        k -= 1
    return k


def global_seeds(n_total, base_seed=17):
    """One Philox key per *global* replica id, independent of how replicas are sharded."""
    return (np.arange(n_total, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(base_seed))


def count_columns(n_proc, n_occ):
    """Columns of a tally row that are event / step / replica *counts* (kmos_b200_reduce_tallies layout:
    procstat[P], integ_rates[P], occupation[n_occ], kmc_time, kmc_steps, n_replicas)."""
    return list(range(n_proc)) + [2 * n_proc + n_occ + 1, 2 * n_proc + n_occ + 2]


def all_reduce_tallies(tally, group=None, count_cols=None):
    """Sum a per-group tally tensor ([n_groups, words] float64, device or CPU) over all ranks, in place.

    count_cols: the columns holding counts (count_columns): they travel as an int64 all-reduce --
    exact for any total -- and the float64 sums (integ_rates, occupation, kmc_time) as a second one.  Without
    it everything travels as float64 (counts stay exact below 2**53)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return tally
    if count_cols is None:
        dist.all_reduce(tally, op=dist.ReduceOp.SUM, group=group)
        return tally
    t2 = tally.view(-1, tally.shape[-1])
    idx = torch.as_tensor(count_cols, dtype=torch.long, device=tally.device)
    counts = t2.index_select(1, idx).round().to(torch.int64)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(tally, op=dist.ReduceOp.SUM, group=group)
    t2.index_copy_(1, idx, counts.to(torch.float64))
    return tally


def merge_group_tallies(tally, group_offset, n_groups_total):
    """Place this rank's [n_local_groups, words] block into a zeroed [n_groups_total, words] array so that a
    plain SUM all-reduce concatenates sweep points that live on different ranks."""
    import torch
    out = torch.zeros((n_groups_total, tally.shape[1]), dtype=tally.dtype, device=tally.device)
    out[group_offset:group_offset + tally.shape[0]] = tally
    return out
