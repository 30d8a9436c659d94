"""C-ABI entry points not covered by the parity suites: rate setters/getters, time setter, tallies,
kernel selection, argument validation."""
import numpy as np
import pytest

from conftest import load_model
from kmos_b200 import capi

pytestmark = pytest.mark.gpu


def test_rate_setters_tallies_and_validation():
    from kmos_b200 import engine
    ir, blob, info = load_model("zgb_local_smart")
    m = engine.Model(ir=ir, blob=blob, info=info)
    R = 6
    rates = np.tile(np.array([0.5, 0.25, 0.25, 1e3, 1e3, 1e3, 1e3, 1e-3, 1e-3, 1e-3]), (R, 1))
    b = engine.Batch(m, R, [10, 10], rates=rates)
    assert np.array_equal(b.rates, rates)
    b.set_rate_const(1, 0.75, replica=2)          # base.set_rate_const(proc, rate) on one replica
    b.set_rate_const(4, 2e3)                       # broadcast
    got = b.rates
    assert got[2, 0] == 0.75 and np.all(got[:, 3] == 2e3) and got[0, 0] == 0.5
    with pytest.raises(capi.KmosB200Error):
        b.set_rates(-rates)                        # negative rate constants are rejected
    with pytest.raises(capi.KmosB200Error):
        b.set_rate_const(99, 1.0)
    b.do_steps(300)
    acc = b.accum_rates                            # base.update_accum_rate + get_accum_rate
    ns = b.nr_of_sites
    np.testing.assert_allclose(acc[:, -1], (ns * got).sum(axis=1), rtol=1e-12)
    # tallies, host path: two groups
    t = b.split_tally(b.reduce_tallies(np.array([0, 0, 0, 1, 1, 1]), 2))
    ps = b.procstat
    assert np.array_equal(t["procstat"][0], ps[:3].sum(axis=0)) and np.array_equal(t["procstat"][1], ps[3:].sum(axis=0))
    np.testing.assert_allclose(t["kmc_time"], [b.kmc_time[:3].sum(), b.kmc_time[3:].sum()], rtol=1e-14)
    np.testing.assert_allclose(t["occupation"][0], b.occupation[:3].sum(axis=0).reshape(-1), rtol=1e-14)
    assert list(t["n_replicas"]) == [3, 3] and list(t["kmc_steps"]) == [900, 900]
    # base.set_kmc_time
    capi.check(b.L.kmos_b200_set_kmc_time(b.h, np.zeros(R)))
    assert np.all(b.kmc_time == 0)
    b.do_steps(10)
    assert np.all(b.kmc_time > 0) and np.all(b.kmc_step == 310)
    b.close()


def test_kernel_selection_and_fallback_reasons():
    from kmos_b200 import engine
    ir, blob, info = load_model("pairwise_lat_int")
    m = engine.Model(ir=ir, blob=blob, info=info)
    b = engine.Batch(m, 2, [8, 8], rates=np.ones((2, 19)))
    assert b.kernel_info()["kernel_name"] == "warp_hbm"
    with pytest.raises(capi.KmosB200Error, match="shared-memory kernel unavailable"):
        b.select_kernel(capi.KERNEL_SMEM)
    b.select_kernel(capi.KERNEL_GENERIC)
    assert b.kernel_info()["kernel_name"] == "generic"
    b.close()
    # a lattice smaller than twice the interaction range cannot use the compact class entries
    ir, blob, info = load_model("zgb_local_smart")
    m = engine.Model(ir=ir, blob=blob, info=info)
    # neighbour offsets alias under the periodic wrap of a 2x2 lattice: only the byte-code engine, which
    # executes the generated statements one by one, is valid there
    b = engine.Batch(m, 2, [2, 2], rates=np.ones((2, 10)))
    assert b.kernel_info()["kernel_name"] == "generic"
    with pytest.raises(capi.KmosB200Error, match="warp-per-replica HBM kernel unavailable"):
        b.select_kernel(capi.KERNEL_WARP_HBM)
    b.do_steps(50)
    assert np.all(b.kmc_step == 50)
    b.close()


@pytest.mark.parametrize("name,size,R", [
    ("mini_101_local_smart", [5, 5], 5000),       # more replicas than one member tile (2048) of the tally kernel
    ("pairwise84_local_smart", [8, 8], 300),      # 2 P + ... > 128 words: more than one word per thread
])
def test_tallies_are_ordered_group_sums(name, size, R):
    """kb_tally_kernel: per-group sums over interleaved groups, added in replica order (so they are reproducible
    bit for bit on the host), an empty group, and the ungrouped call."""
    from kmos_b200 import engine
    from util import make_inputs
    ir, blob, info = load_model(name)
    m = engine.Model(ir=ir, blob=blob, info=info)
    rates, lut, seeds = make_inputs(ir, info, R, seed=5)
    b = engine.Batch(m, R, size, seeds=seeds, rates=rates)
    b.do_steps(200)
    groups = (np.arange(R) * 7 % 5).astype(np.int32)   # five interleaved groups, a sixth stays empty
    t = b.split_tally(b.reduce_tallies(groups, 6))
    ps, integ, occ, time, steps = b.procstat, b.integ_rates, b.occupation.reshape(R, -1), b.kmc_time, b.kmc_step

    def ordered(values, members):
        acc = np.zeros(values.shape[1:])
        for r in members:
            acc = acc + values[r]
        return acc
    for g in range(6):
        members = np.nonzero(groups == g)[0]
        assert t["n_replicas"][g] == len(members) and t["kmc_steps"][g] == steps[members].sum()
        assert np.array_equal(t["procstat"][g], ps[members].sum(axis=0))
        assert np.array_equal(t["integ_rates"][g], ordered(integ, members))
        assert np.array_equal(t["occupation"][g], ordered(occ, members))
        assert t["kmc_time"][g] == ordered(time[:, None], members)[0]
    assert t["n_replicas"][5] == 0 and not t["procstat"][5].any()
    one = b.split_tally(b.reduce_tallies())
    assert one["n_replicas"][0] == R and np.array_equal(one["procstat"][0], ps.sum(axis=0))
    assert one["kmc_time"][0] == ordered(time[:, None], range(R))[0]
    b.close()
