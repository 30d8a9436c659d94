"""One rank of tests/test_gpu_multi.py (launched by torch.distributed.run, NCCL): steps its shard of the
replicas, reduces the tallies over the ranks, and rank 0 compares with the same replicas on one GPU."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    from kmos_b200 import engine, parallel, tables, workloads
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    name, size, R_total, n, n_groups = "ruo2_local_smart", [20, 20], 600, 1500, 6
    ir = tables.load_ir(os.path.join(os.path.dirname(HERE), "tests", "golden", "models", name + ".json"))
    model = engine.Model(ir=ir)
    P = model.n_proc
    rates_all = workloads.rates_for("ruo2", ir, R_total)
    seeds_all = parallel.global_seeds(R_total)
    groups_all = (np.arange(R_total) * n_groups // R_total).astype(np.int32)  # groups straddle the ranks
    lo, hi = parallel.shard_bounds(R_total, rank, world)
    assert all(parallel.shard_of_replica(R_total, world, r) == rank for r in (lo, hi - 1))
    b = engine.Batch(model, hi - lo, size, device=local, seeds=seeds_all[lo:hi],
                     replica_ids=np.arange(lo, hi, dtype=np.uint32), rates=rates_all[lo:hi])
    b.do_steps(n)
    words = b.tally_words()
    tally = torch.zeros((n_groups, words), dtype=torch.float64, device="cuda")
    b.reduce_tallies(groups_all[lo:hi], n_groups, dev_ptr=tally.data_ptr(), want_host=False)
    torch.cuda.synchronize()
    cols = parallel.count_columns(P, model.n_species * model.spuck)
    parallel.all_reduce_tallies(tally, count_cols=cols)
    got = tally.cpu().numpy()
    b.close()
    if rank == 0:
        one = engine.Batch(model, R_total, size, device=local, seeds=seeds_all,
                           replica_ids=np.arange(R_total, dtype=np.uint32), rates=rates_all)
        one.do_steps(n)
        ref = one.reduce_tallies(groups_all, n_groups)
        t_got, t_ref = one.split_tally(got), one.split_tally(ref)
        for k in ("procstat", "kmc_steps", "n_replicas"):
            assert np.array_equal(t_got[k], t_ref[k]), k          # counts: exact (int64 all-reduce)
        assert t_got["n_replicas"].sum() == R_total and t_got["kmc_steps"].sum() == R_total * n
        for k in ("integ_rates", "occupation", "kmc_time"):
            np.testing.assert_allclose(t_got[k], t_ref[k], rtol=1e-12, atol=0, err_msg=k)
        one.close()
        print("MULTI_GPU_OK world=%d" % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
