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
    assert b.kernel_info()["kernel_name"] in ("generated", "smem", "warp_hbm")  # planner's choice; all parity-tested
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
    assert b.kernel_info()["kernel_name"] == "warp_hbm"
    b.do_steps(n)
    _check_sample(b, blob, [256, 256], seeds, rates, lut, n, [0, 7], avail=True)


@pytest.mark.parametrize("lanes", ["lanes", "lane0"])
def test_config4_on_the_production_kernel_256x256(lanes, monkeypatch):
    """Config E's lattice on kb_otf_fast.cuh: 65 536 cells, i.e. the 32-bit list instantiation the bench runs and
    256 blocks per row -- the small lattices of tests/test_gpu_otf_fast.py use 16-bit lists.  Over 400 steps from
    the empty lattice the selection coincides with the exact one: lattice, counts and lists bit for bit."""
    from kmos_b200 import capi, engine
    from oracle import oracle
    monkeypatch.setenv("KMOS_B200_OTF_LANES", "1" if lanes == "lanes" else "0")
    ir, blob, info = load_model("pairwise_otf_otf")
    R, n = 6, 400
    rates = workloads.rates_for("pairwise_otf", ir, R)
    lut = np.stack([otf_mod.build_lut(ir, info, rates[r]) for r in range(R)])
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(17)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, [256, 256], seeds=seeds, rates=rates, lut=lut,
                     kernel=capi.KERNEL_OTF_FAST)
    assert b.kernel_info()["kernel_name"] == "otf_fast"
    b.do_steps(n)
    assert np.all(b.status == 0) and np.all(b.kmc_step == n) and np.all(b.procstat.sum(axis=1) == n)
    for r in (0, R - 1):
        o = oracle.Oracle(blob, [256, 256], seed=int(seeds[r]), replica=r, rates=rates[r], lut=lut[r])
        assert o.do_steps(n) == 0
        assert np.array_equal(b.lattice[r], o.lattice) and np.array_equal(b.procstat[r], o.procstat)
        assert np.array_equal(b.nr_of_sites[r], o.nr_of_sites)
        assert np.array_equal(b.avail_sites(r), o.avail_sites)
        assert abs(b.kmc_time[r] - o.kmc_time) <= 1e-9 * o.kmc_time
    b.close()


def test_production_stream_statistics_within_3_sigma():
    """north_star: production-stream runs must agree statistically with the reference.  GPU ensemble (one set
    of Philox keys) vs CPU-oracle ensemble (a disjoint set of keys) of the ZGB model at y_CO = 0.45: the mean
    coverages and the CO2 turn-over frequency agree within 3 standard errors."""
    from kmos_b200 import engine, rates as rates_mod
    from oracle import oracle
    ir, blob, info = load_model("zgb_local_smart")
    size, warm, n = [16, 16], 20000, 20000
    r = np.asarray(rates_mod.model_rates(ir, {"yCO": 0.45}))
    ox = [i for i, p in enumerate(ir["procs"]) if p.lower().startswith("co_oxidation")]

    def observables(occ, ps0, ps1, t0, t1):
        tof = (ps1 - ps0)[:, ox].sum(axis=1) / (t1 - t0) / (size[0] * size[1])
        return np.column_stack([occ.reshape(len(occ), -1), tof])

    R = 96
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(1000)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, size, seeds=seeds, rates=np.tile(r, (R, 1)))
    b.do_steps(warm)
    ps0, t0 = b.procstat, b.kmc_time
    b.do_steps(n)
    gpu = observables(b.occupation, ps0, b.procstat, t0, b.kmc_time)

    Rc = 48
    cpu_rows = []
    for k in range(Rc):
        o = oracle.Oracle(blob, size, seed=500000 + k, replica=k, rates=r)
        o.do_steps(warm)
        p0, tt0 = o.procstat.copy(), o.kmc_time
        o.do_steps(n)
        cpu_rows.append(observables(o.occupation[None], p0[None], o.procstat[None], np.array([tt0]),
                                    np.array([o.kmc_time]))[0])
    cpu = np.asarray(cpu_rows)
    diff = gpu.mean(axis=0) - cpu.mean(axis=0)
    sigma = np.sqrt(gpu.var(axis=0, ddof=1) / R + cpu.var(axis=0, ddof=1) / Rc)
    ok = (np.abs(diff) <= 3 * sigma) | (sigma == 0)
    assert ok.all(), (diff, sigma)
    assert gpu[:, -1].mean() > 0  # the reactive window: CO2 is produced


def test_long_validation_ruo2_256_replicas_1e6_steps():
    """SURVEY 8d correctness gate: >= 256 replicas, >= 1e6 kMC steps each, checked every 1e5 steps -- lattice and
    procstat bit-exact, kmc_time within 1e-12 relative -- on the headline configuration (RuO2 20x20, sweep
    points spread over the T x p_CO grid), against the oracle on all host cores."""
    import os
    from kmos_b200 import engine
    from util import oracle_checkpoints
    ir, blob, info = load_model("ruo2_local_smart")
    R, chunks = 256, [100000] * 10
    grid = workloads.rates_for("ruo2", ir, 16384)
    rates = np.ascontiguousarray(grid[:: 16384 // R][:R])
    seeds = np.arange(R, dtype=np.uint64) * np.uint64(104729) + np.uint64(17)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, [20, 20], seeds=seeds, rates=rates)
    assert b.kernel_info()["kernel_name"] == "generated"
    got = []
    for n in chunks:
        b.do_steps(n)
        got.append((b.lattice.astype(np.int8), b.procstat, b.kmc_time, b.kmc_step, b.status))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    olat, ops, ot, ostep, ost = oracle_checkpoints(blob, [20, 20], seeds, rates, chunks, cores)
    for c in range(len(chunks)):
        lat, ps, t, step, st = got[c]
        assert np.all(st == 0) and np.all(ost[:, c] == 0)
        assert np.all(step == sum(chunks[: c + 1])) and np.all(ostep[:, c] == step)
        bad = np.nonzero((lat != olat[:, c]).any(axis=1))[0]
        assert bad.size == 0, "lattice differs, checkpoint %d, replicas %s" % (c, bad[:8])
        assert np.array_equal(ps, ops[:, c]), "procstat differs, checkpoint %d" % c
        np.testing.assert_allclose(t, ot[:, c], rtol=1e-12, atol=0)
