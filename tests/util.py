"""Shared helpers for the parity tests."""
import numpy as np

from conftest import load_model
from kmos_b200 import otf as otf_mod
from oracle import oracle


def make_inputs(ir, info, R, seed=0, spread=1.0):
    """Deterministic per-replica rate constants (log-uniform around 1) and otf tables."""
    rng = np.random.RandomState(seed)
    P = len(ir["procs"])
    rates = np.exp(rng.uniform(-spread, spread, (R, P)))
    lut = None
    if ir["backend"] == "otf":
        lut = np.stack([otf_mod.build_lut(ir, info, rates[r]) for r in range(R)])
    seeds = (np.arange(R, dtype=np.uint64) * np.uint64(7919) + np.uint64(12345))
    return rates, lut, seeds


def run_oracles(blob, size, rates, lut, seeds, chunks, replica_ids=None):
    """One Oracle per replica; yields after each chunk the list of oracles."""
    R = rates.shape[0]
    os_ = [oracle.Oracle(blob, size, seed=int(seeds[r]), replica=int(r if replica_ids is None else replica_ids[r]),
                         rates=rates[r], lut=None if lut is None else lut[r]) for r in range(R)]
    yield os_
    for n in chunks:
        for o in os_:
            o.do_steps(n)
        yield os_


def compare_batch(batch, oracles, avail_replicas=(0,), time_rtol=1e-12, integ_rtol=1e-10):
    R = len(oracles)
    st = batch.status
    assert np.array_equal(st, np.array([o.status[0] for o in oracles])), st
    assert np.array_equal(batch.kmc_step, np.array([o.kmc_step for o in oracles]))
    assert np.array_equal(batch.lattice, np.stack([o.lattice for o in oracles])), "lattice differs"
    assert np.array_equal(batch.procstat, np.stack([o.procstat for o in oracles])), "procstat differs"
    assert np.array_equal(batch.nr_of_sites, np.stack([o.nr_of_sites for o in oracles])), "nr_of_sites differs"
    t_ref = np.array([o.kmc_time for o in oracles])
    np.testing.assert_allclose(batch.kmc_time, t_ref, rtol=time_rtol, atol=0)
    np.testing.assert_allclose(batch.integ_rates, np.stack([o.integ_rates for o in oracles]), rtol=integ_rtol, atol=0)
    np.testing.assert_allclose(batch.occupation, np.stack([o.occupation for o in oracles]), rtol=0, atol=1e-15)
    for r in avail_replicas:
        if r < R:
            assert np.array_equal(batch.avail_sites(r), oracles[r].avail_sites), "avail_sites differ (replica %d)" % r


def oracle_checkpoints(blob, size, seeds, rates, chunks, workers, timeout=600, lut=None):
    """Oracle trajectories of len(seeds) replicas on `workers` host processes (tests/oracle_worker.py started
    with subprocess: independent of this process' CUDA context, bounded by `timeout` seconds).
    -> (lattice[R][C][V] int8, procstat[R][C][P], kmc_time[R][C], kmc_step[R][C], status[R][C])."""
    import os
    import subprocess
    import sys
    import tempfile
    R = len(seeds)
    workers = max(1, min(int(workers), R))
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for w in range(workers):
            ids = np.arange(w, R, workers)
            fin, fout = os.path.join(tmp, "in%d.npz" % w), os.path.join(tmp, "out%d.npz" % w)
            extra = {} if lut is None else {"lut": np.asarray(lut)[ids]}
            np.savez(fin, blob=blob, size=np.asarray(size), seeds=np.asarray(seeds)[ids], ids=ids,
                     rates=np.asarray(rates)[ids], chunks=np.asarray(chunks), **extra)
            procs.append((ids, fout, subprocess.Popen([sys.executable, os.path.join(here, "oracle_worker.py"), fin, fout])))
        out = None
        for ids, fout, p in procs:
            try:
                rc = p.wait(timeout=timeout)
            except subprocess.TimeoutExpired:
                p.kill()
                raise
            assert rc == 0, "oracle worker failed"
            d = np.load(fout)
            if out is None:
                out = [np.zeros((R,) + d[k].shape[1:], d[k].dtype) for k in ("lattice", "procstat", "time", "step", "status")]
            for a, k in zip(out, ("lattice", "procstat", "time", "step", "status")):
                a[ids] = d[k]
    return out
