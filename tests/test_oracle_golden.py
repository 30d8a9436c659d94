"""Pin the CPU oracle against the reference's own known-answer trajectory.

Reference test: tests/test_run/test_run.py:40-70 -- AB model, 20x20, seed 1, 10000 x
(get_next_kmc_step, run_proc_nr); the (proc, site) list must equal ref_procs_sites_<backend>.log.
That log is produced under gfortran's random_number (xoshiro256**, random_seed(put=)), which the
oracle restates (oracle/kmos_oracle.c: gfortran_seed / gfortran_random_r8).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_model
from kmos_b200 import rates
from oracle import oracle


def _replay(backend, n):
    ir, blob, info = load_model("ab_" + backend, with_device=False)
    r = rates.model_rates(ir)
    lut = None
    if backend == "otf":
        lut = np.zeros(info["lut_total"])
        for g in info["gr"].values():
            lut[g["lut_offset"]] = r[g["proc"] - 1]
    o = oracle.Oracle(blob, [20, 20], seed=1, rng=oracle.RNG_GFORTRAN, rates=r, lut=lut)
    out = np.zeros((n, 2), dtype=np.int32)
    lattices = []
    for i in range(n):
        p, s, st = o.get_next_kmc_step()
        assert st == oracle.OK
        out[i] = (p, s)
        o.run_proc_nr(p, s)
        lattices.append(o.lattice)
    return out, lattices


def test_local_smart_reproduces_reference_golden():
    ref = np.load(os.path.join(GOLDEN, "ab_ref_procs_sites.npy"))
    assert ref.shape == (10000, 2)
    got, _ = _replay("local_smart", 10000)
    mism = np.nonzero((got != ref).any(axis=1))[0]
    assert mism.size == 0, "first mismatch at event %d: got %s, reference %s" % (
        mism[0], got[mism[0]], ref[mism[0]])


@pytest.mark.parametrize("backend", ["lat_int", "otf"])
def test_other_backends_agree_physically_with_golden_prefix(backend):
    """The upstream lat_int/otf logs are copies of the local_smart run (module caching in test_run.py),
    so they cannot pin those backends.  What can be checked: under the same RNG stream the lat_int / otf
    restatements walk through the same lattice configurations as the pinned local_smart one until the
    different avail_sites ordering of those generators makes the trajectories part (>= 150 events)."""
    _, ref_l = _replay("local_smart", 150)
    _, got_l = _replay(backend, 150)
    for i, (a, b) in enumerate(zip(ref_l, got_l)):
        assert np.array_equal(a, b), "lattice differs after event %d" % i
