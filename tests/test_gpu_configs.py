"""The five BASELINE.json configurations at their full sizes: size-independent invariants on every replica
plus bit-exact oracle parity on a sample of replicas (the oracle finishes those in seconds)."""
import numpy as np
import pytest

from conftest import load_model
from kmos_b200 import otf as otf_mod, workloads

pytestmark = pytest.mark.gpu


def _check_sample(batch, blob, size, seeds, rates, lut, n, sample, avail=True):
    from oracle import oracle
    lat, ps, ns, t = batch.lattice, batch.procstat, batch.nr_of_sites, batch.kmc_time
    assert np.all(batch.status == 0)
    assert np.all(batch.kmc_step == n)
    assert np.all(ps.sum(axis=1) == n)
    for r in sample:
        o = oracle.Oracle(blob, size, seed=int(seeds[r]), replica=r, rates=rates[r], lut=None if lut is None else lut[r])
        assert o.do_steps(n) == 0
        assert np.array_equal(lat[r], o.lattice), "lattice, replica %d" % r
        assert np.array_equal(ps[r], o.procstat), "procstat, replica %d" % r
        assert np.array_equal(ns[r], o.nr_of_sites), "nr_of_sites, replica %d" % r
        assert abs(t[r] - o.kmc_time) <= 1e-12 * o.kmc_time
        if avail:
            assert np.array_equal(batch.avail_sites(r), o.avail_sites), "avail_sites, replica %d" % r


def test_config0_mini_101_kmos_benchmark_1e6_steps():
    """`kmos benchmark` (kmos/cli.py:293-312): mini_101, 20x20, 1e6 steps."""
    from kmos_b200 import engine
    ir, blob, info = load_model("mini_101_local_smart")
    R, n = 4, 1000000
    rates = workloads.rates_for("mini_101", ir, R)
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(1)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, [20, 20], seeds=seeds, rates=rates)
    b.do_steps(n)
    _check_sample(b, blob, [20, 20], seeds, rates, None, n, [0, 3])


def test_config1_zgb_64x64_4096_replicas():
    from kmos_b200 import engine
    ir, blob, info = load_model("zgb_local_smart")
    R, n = 4096, 400
    rates, group_of, grid = workloads.zgb_grid(ir)
    assert rates.shape[0] == R
    seeds = np.arange(R, dtype=np.uint64) * np.uint64(3) + np.uint64(7)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, [64, 64], seeds=seeds, rates=rates)
    assert b.kernel_info()["kernel_name"] == "smem"
    b.do_steps(n)
    _check_sample(b, blob, [64, 64], seeds, rates, None, n, [0, 2047, 4095])
    np.testing.assert_allclose(b.occupation.sum(axis=1), 1.0, atol=1e-12)


def test_config3_pairwise_lat_int_128x128_2048_replicas():
    from kmos_b200 import engine
    ir, blob, info = load_model("pairwise_lat_int")
    R, n = 2048, 150
    rates = workloads.rates_for("pairwise", ir, R) * (1.0 + 0.001 * (np.arange(R) % 7))[:, None]
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(99)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, [128, 128], seeds=seeds, rates=rates)
    assert b.kernel_info()["kernel_name"] == "warp_hbm"
    b.do_steps(n)
    _check_sample(b, blob, [128, 128], seeds, rates, None, n, [0, 2047], avail=True)


def test_config4_pairwise_otf_256x256():
    from kmos_b200 import engine
    ir, blob, info = load_model("pairwise_otf_otf")
    R, n = 8, 120
    rates = workloads.rates_for("pairwise_otf", ir, R)
    lut = np.stack([otf_mod.build_lut(ir, info, rates[r]) for r in range(R)])
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(5)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, [256, 256], seeds=seeds, rates=rates, lut=lut)
    assert b.kernel_info()["kernel_name"] == "generic"
    b.do_steps(n)
    _check_sample(b, blob, [256, 256], seeds, rates, lut, n, [0, 7], avail=True)
