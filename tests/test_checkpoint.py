"""Restart files in the reference's .reload format (kmos_b200/checkpoint.py; base.mpy:365-578)."""
import numpy as np
import pytest

from conftest import load_model
from kmos_b200 import checkpoint


def _oracle_state(o, rates):
    return dict(kmc_time=o.kmc_time, kmc_step=o.kmc_step, procstat=o.procstat, nr_of_sites=o.nr_of_sites,
                rates=rates, integ_rates=o.integ_rates, lattice=o.lattice, avail_sites=o.avail_sites)


def test_reload_file_round_trip_and_layout(tmp_path):
    from oracle import oracle
    ir, blob, info = load_model("ab_local_smart")
    rates = np.linspace(0.5, 2.0, len(ir["procs"]))
    o = oracle.Oracle(blob, [6, 5], seed=9, replica=0, rates=rates)
    o.do_steps(500)
    path = str(tmp_path / "ab.reload")
    st = _oracle_state(o, rates)
    checkpoint.write_reload(path, st)
    lines = open(path).read().splitlines()
    # the reference's header and labels, in its order (base.mpy:524-566)
    assert lines[0] == "#Reload file written by kmos. Do not edit manually!"
    labels = [ln.split()[0] for ln in lines if not ln.startswith("#")]
    P = len(ir["procs"])
    assert labels[:5] == ["kmc_time", "walltime", "kmc_step", "nr_of_proc", "volume"]
    assert labels.count("avail_sites") == P and labels.count("avail_sites_back") == P
    back = checkpoint.read_reload(path)
    assert back["kmc_step"] == 500 and back["kmc_time"] == pytest.approx(o.kmc_time, rel=1e-15)
    for k in ("procstat", "nr_of_sites", "lattice", "avail_sites"):
        assert np.array_equal(back[k], np.asarray(st[k]).reshape(back[k].shape)), k
    assert np.array_equal(back["integ_rates"], o.integ_rates)  # hex floats: exact
    np.testing.assert_allclose(back["rates"], rates, rtol=1e-7)   # the reference's es14.7


@pytest.mark.gpu
def test_resume_from_reload_file_is_bit_identical(tmp_path):
    from kmos_b200 import engine
    ir, blob, info = load_model("ruo2_local_smart")
    from kmos_b200 import workloads
    R = 6
    rates = workloads.rates_for("ruo2", ir, 16384)[:: 16384 // R][:R].copy()
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(77)
    model = engine.Model(ir=ir, blob=blob, info=info)
    a = engine.Batch(model, R, [20, 20], seeds=seeds, rates=rates)
    a.do_steps(3000)
    paths = []
    for r in range(R):
        paths.append(str(tmp_path / ("rep%d.reload" % r)))
        a.save_system(paths[-1], r)
    a.do_steps(5000)
    b = engine.Batch(model, R, [20, 20], seeds=seeds, rates=rates)
    for r in range(R):
        b.reload_system(paths[r], r)
    assert np.all(b.kmc_step == 3000)
    b.do_steps(5000)
    assert np.array_equal(a.lattice, b.lattice)
    assert np.array_equal(a.procstat, b.procstat)
    assert np.array_equal(a.kmc_time, b.kmc_time)
    assert np.array_equal(a.integ_rates, b.integ_rates)
    for r in (0, R - 1):
        assert np.array_equal(a.avail_sites(r), b.avail_sites(r))
    # a corrupted file is rejected, not loaded
    st = checkpoint.read_reload(paths[0])
    q = int(np.argmax(st["nr_of_sites"]))  # a process that has sites: break the link between its two planes
    st["avail_sites"][q, 0, 0] = st["avail_sites"][q, 0, 0] % st["volume"] + 1
    st["rates"] = rates[0]
    bad = str(tmp_path / "bad.reload")
    checkpoint.write_reload(bad, st)
    with pytest.raises(Exception):
        b.reload_system(bad, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name,size", [("intzgb_otf", [10, 9]), ("pairwise_otf_otf", [16, 16])])
def test_otf_resume_from_reload_file_is_bit_identical(tmp_path, name, size):
    """otf restart: the reference's file has no rates_matrix (base_otf.f90:602-663); reload
    rebuilds every row from the restored lattice and the gr_<proc> table, and the resumed batch must walk the
    same trajectory as the uninterrupted one and as the oracle."""
    from kmos_b200 import engine
    from util import make_inputs, run_oracles, compare_batch
    ir, blob, info = load_model(name)
    R = 4
    rates, lut, seeds = make_inputs(ir, info, R, seed=3)
    model = engine.Model(ir=ir, blob=blob, info=info)
    a = engine.Batch(model, R, size, seeds=seeds, rates=rates, lut=lut)
    a.do_steps(1500)
    paths = []
    for r in range(R):
        paths.append(str(tmp_path / ("rep%d.reload" % r)))
        a.save_system(paths[-1], r)
    a.do_steps(2500)
    b = engine.Batch(model, R, size, seeds=seeds, rates=rates, lut=lut)
    for r in range(R):
        b.reload_system(paths[r], r)
    assert np.all(b.kmc_step == 1500)
    b.do_steps(2500)
    assert np.array_equal(a.lattice, b.lattice) and np.array_equal(a.procstat, b.procstat)
    assert np.array_equal(a.kmc_time, b.kmc_time)
    for r in (0, R - 1):
        assert np.array_equal(a.avail_sites(r), b.avail_sites(r))
    gen = run_oracles(blob, size, rates, lut, seeds, [4000])
    next(gen)
    compare_batch(b, next(gen), avail_replicas=(0, R - 1))


@pytest.mark.gpu
def test_otf_parameter_change_mid_run_refreshes_registered_rates():
    """set_otf_lut (KMC_Model.set_parameters on an otf model) must refresh the rates of events that
    are already registered -- the reference's set_rate_constants ends with proclist.recalculate_rates_matrix
    (proclist_generic_subroutines.mpy:307-325)."""
    from kmos_b200 import capi, engine, otf as otf_mod
    from util import make_inputs, compare_batch
    from oracle import oracle
    for name, size in (("intzgb_otf", [10, 9]), ("pairwise_otf_otf", [12, 12])):
        ir, blob, info = load_model(name)
        R = 4
        rates, lut, seeds = make_inputs(ir, info, R, seed=8)
        rates2 = rates * np.exp(np.random.RandomState(2).uniform(-0.5, 0.5, rates.shape))
        lut2 = np.stack([otf_mod.build_lut(ir, info, rates2[r]) for r in range(R)])
        for kernel in (capi.KERNEL_WARP_HBM, capi.KERNEL_GENERIC):
            b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, size, seeds=seeds, rates=rates, lut=lut,
                             kernel=kernel)
            os_ = [oracle.Oracle(blob, size, seed=int(seeds[r]), replica=r, rates=rates[r], lut=lut[r]) for r in range(R)]
            b.do_steps(1200)
            b.set_rates(rates2)
            b.set_otf_lut(lut2)
            for r, o in enumerate(os_):
                o.do_steps(1200)
                o.set_rates(rates2[r])
                o.set_lut(lut2[r])
                o.recalculate_rates_matrix()
            b.do_steps(1800)
            for o in os_:
                o.do_steps(1800)
            compare_batch(b, os_, avail_replicas=(0, R - 1))
            b.close()
