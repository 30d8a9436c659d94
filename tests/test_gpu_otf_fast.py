"""otf production kernel (kb_otf_fast.cuh, KMOS_B200_KERNEL_OTF_FAST): sub-linear event selection over block sums
of rates_matrix -- the O(log N)-class replacement the reference's documentation announces for its O(N_sites)
scheme (doc/source/topic_guides/otf_backend.rst:198-202).  It keeps the reference's prefix order, so it selects
the reference's (process, site) unless a random number falls within rounding distance of an interval boundary:
over a few thousand steps the trajectory must coincide with the oracle's (lattice, procstat, avail_sites
bit-exact; kmc_time to 1e-9 because the row totals are associated differently), and over long runs ensembles
must agree statistically with the exact kernel (3 sigma on coverages and process rates)."""
import numpy as np
import pytest

from conftest import load_model
from kmos_b200 import capi
from util import make_inputs, run_oracles

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,size,chunks", [
    ("pairwise_otf_otf", [24, 20], [1500, 1500]),        # 2 blocks per row
    ("pairwise_otf_otf", [64, 48], [2100, 900]),         # 12 blocks per row, crosses a re-accumulation (2048)
    ("intzgb_otf", [20, 18], [1500, 1500]),
    ("ruo2default_otf", [20, 20], [1200, 1200]),         # 36 processes, 2 sites per cell
    ("multidentate_otf", [20, 18], [1500, 1500]),        # species spanning two and four sites
    ("hop3d_otf", [8, 7, 6], [1500, 1500]),              # z offsets
    ("ab_otf", [20, 20], [1500, 1500]),
    ("zgb_otf", [24, 22], [1500, 1500]),
    ("pt111_otf", [20, 18], [1500, 1500]),               # two hollow sites per cell
])
@pytest.mark.parametrize("lanes", ["lanes", "lanes_routine_tail", "lane0"])
def test_fast_selection_walks_the_exact_trajectory(name, size, chunks, lanes, monkeypatch):
    """lanes: the event's guarded dels, lattice writes, guarded rate updates and the flattened if-tree of add_proc
    statements spread over the warp (devtables.compile_otf_tables); lanes_routine_tail: the add block kept as a
    byte-code routine run by lane 0 (what very large blocks fall back to); lane0: the whole routine interpreted
    by lane 0."""
    from kmos_b200 import engine
    monkeypatch.setenv("KMOS_B200_OTF_LANES", "0" if lanes == "lane0" else "1")
    monkeypatch.setenv("KMOS_B200_OTF_FLATTEN", "0" if lanes == "lanes_routine_tail" else "1")
    ir, blob, info = load_model(name)
    assert info["device"]["supported"], info["device"].get("reason")
    assert (info["device"]["routine_tails"] > 0) == (lanes == "lanes_routine_tail")
    R = 5
    rates, lut, seeds = make_inputs(ir, info, R, seed=len(name) + 3)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, size, seeds=seeds, rates=rates, lut=lut,
                     kernel=capi.KERNEL_OTF_FAST)
    assert b.kernel_info()["kernel_name"] == "otf_fast"
    gen = run_oracles(blob, size, rates, lut, seeds, chunks)
    next(gen)
    for n, oracles in zip(chunks, gen):
        b.do_steps(n)
        assert np.array_equal(b.status, np.zeros(R, np.int32))
        assert np.array_equal(b.kmc_step, np.array([o.kmc_step for o in oracles]))
        assert np.array_equal(b.lattice, np.stack([o.lattice for o in oracles])), "lattice differs"
        assert np.array_equal(b.procstat, np.stack([o.procstat for o in oracles])), "procstat differs"
        assert np.array_equal(b.nr_of_sites, np.stack([o.nr_of_sites for o in oracles]))
        for r in (0, R - 1):
            assert np.array_equal(b.avail_sites(r), oracles[r].avail_sites)
        np.testing.assert_allclose(b.kmc_time, [o.kmc_time for o in oracles], rtol=1e-9)
        np.testing.assert_allclose(b.integ_rates, np.stack([o.integ_rates for o in oracles]), rtol=1e-8)
    # switching back to the exact kernel continues bit-exactly from the shared state (row totals are re-added)
    b.select_kernel(capi.KERNEL_WARP_HBM)
    b.do_steps(500)
    for o in oracles:
        o.do_steps(500)
    assert np.array_equal(b.lattice, np.stack([o.lattice for o in oracles]))
    assert np.array_equal(b.procstat, np.stack([o.procstat for o in oracles]))
    b.close()


def test_fast_selection_is_refused_where_it_cannot_work():
    from kmos_b200 import engine
    ir, blob, info = load_model("ruo2default_otf")   # 36 processes on 4 x 4 cells: no room for the block sums
    rates, lut, seeds = make_inputs(ir, info, 2, seed=1)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), 2, [4, 4], seeds=seeds, rates=rates, lut=lut)
    with pytest.raises(capi.KmosB200Error):
        b.select_kernel(capi.KERNEL_OTF_FAST)
    ir2, blob2, info2 = load_model("ab_local_smart")
    r2, _l, s2 = make_inputs(ir2, info2, 2, seed=1)
    b2 = engine.Batch(engine.Model(ir=ir2, blob=blob2, info=info2), 2, [6, 6], seeds=s2, rates=r2)
    with pytest.raises(capi.KmosB200Error):
        b2.select_kernel(capi.KERNEL_OTF_FAST)


def test_fast_selection_agrees_statistically_with_the_exact_kernel():
    """Production-stream check (north_star): ensembles on disjoint Philox keys, exact kernel vs fast kernel,
    coverages and per-process event rates within 3 sigma."""
    from kmos_b200 import engine
    ir, blob, info = load_model("pairwise_otf_otf")
    R, size, warm, n = 96, [32, 32], 6000, 12000
    rates, lut, _seeds = make_inputs(ir, info, 1, seed=4)
    rates, lut = np.tile(rates, (R, 1)), np.tile(lut, (R, 1))
    model = engine.Model(ir=ir, blob=blob, info=info)

    def observables(kernel, seed0):
        seeds = np.arange(R, dtype=np.uint64) * np.uint64(7919) + np.uint64(seed0)
        b = engine.Batch(model, R, size, seeds=seeds, rates=rates, lut=lut, kernel=kernel)
        b.do_steps(warm)
        p0, t0 = b.procstat.astype(float), b.kmc_time
        b.do_steps(n)
        assert np.all(b.status == 0)
        occ = b.occupation.reshape(R, -1)
        ev = (b.procstat - p0) / (b.kmc_time - t0)[:, None]
        b.close()
        return np.hstack([occ, ev])

    exact = observables(capi.KERNEL_WARP_HBM, 1000)
    fast = observables(capi.KERNEL_OTF_FAST, 5000000)
    diff = fast.mean(axis=0) - exact.mean(axis=0)
    sigma = np.sqrt(fast.var(axis=0, ddof=1) / R + exact.var(axis=0, ddof=1) / R)
    ok = (np.abs(diff) <= 3 * sigma) | (sigma == 0)
    assert ok.all(), (diff[~ok], sigma[~ok])
