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
    return min(world - 1, (r * world) // n_replicas) if n_replicas >= world else r % world


def global_seeds(n_total, base_seed=17):
    """One Philox key per *global* replica id, independent of how replicas are sharded."""
    return (np.arange(n_total, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(base_seed))


def all_reduce_tallies(tally, group=None):
    """Sum a per-group tally tensor ([n_groups, words] float64, device or CPU) over all ranks, in place.
    Counts (procstat, kmc_steps, n_replicas) are integers stored in doubles: exact below 2**53."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tally, op=dist.ReduceOp.SUM, group=group)
    return tally


def merge_group_tallies(tally, group_offset, n_groups_total):
    """Place this rank's [n_local_groups, words] block into a zeroed [n_groups_total, words] array so that a
    plain SUM all-reduce concatenates sweep points that live on different ranks."""
    import torch
    out = torch.zeros((n_groups_total, tally.shape[1]), dtype=tally.dtype, device=tally.device)
    out[group_offset:group_offset + tally.shape[0]] = tally
    return out
