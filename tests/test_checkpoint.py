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
