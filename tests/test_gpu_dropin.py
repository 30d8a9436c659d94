"""The kmc_model drop-in package (kmos_b200/dropin/kmc_model) on CUDA: the f2py-shaped calls kmos.run.KMC_Model
makes -- in the order it makes them (kmos/run/__init__.py:243-340, 416-432, 1243-1457) -- against the oracle
under the same Philox stream.  tests/test_dropin_reference_run.py drives the very same package from the
unmodified reference front-end (container only: the reference is not on the GPU box)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REPO, load_model
from kmos_b200 import rates as rates_mod
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def kmc_model(monkeypatch, request):
    name = getattr(request, "param", "ab_local_smart")
    monkeypatch.setenv("KMOS_B200_MODEL", os.path.join(GOLDEN, "models", name + ".json"))
    monkeypatch.syspath_prepend(os.path.join(REPO, "kmos_b200", "dropin"))
    for name in [m for m in sys.modules if m == "kmc_model" or m.startswith("kmc_model.")]:
        del sys.modules[name]
    import kmc_model as km
    yield km
    km.lattice.deallocate_system()
    for name in [m for m in sys.modules if m == "kmc_model" or m.startswith("kmc_model.")]:
        del sys.modules[name]


def test_f2py_shaped_calls_on_cuda_match_the_oracle(kmc_model):
    base, lattice, proclist = kmc_model.base, kmc_model.lattice, kmc_model.proclist
    ir, blob, _info = load_model("ab_local_smart")
    r = np.asarray(rates_mod.model_rates(ir))
    P = len(ir["procs"])
    assert proclist.nr_of_proc == P and lattice.model_dimension == 2 and lattice.spuck == 1
    assert proclist.ab_react_down == ir["procs"].index("AB_react_down") + 1 and proclist.empty == ir["species"].index("empty")
    with pytest.raises(ImportError):
        from kmc_model import proclist_pars  # noqa: F401 -- otf models only, as with the f2py build

    # KMC_Model.__init__ / reset
    proclist.init([20, 20], "kmc_model", lattice.default_layer, 42, True)
    assert base.is_allocated() and list(lattice.system_size) == [20, 20, 1]
    for i in range(P):
        base.set_rate_const(i + 1, float(r[i]))
    base.update_accum_rate()
    o = oracle.Oracle(blob, [20, 20], seed=42, replica=0, rates=r)
    assert [base.get_rate(i + 1) for i in range(P)] == list(r)

    # the loop of the reference's tests/test_run/test_run.py:46-53
    for _ in range(300):
        proc, site = proclist.get_next_kmc_step()
        op_, os_, st = o.get_next_kmc_step()
        assert st == 0 and (int(proc), int(site)) == (op_, os_)
        proclist.run_proc_nr(proc, site)
        o.run_proc_nr(op_, os_)
    # KMC_Model.do_steps + what get_atoms(geometry=False) reads
    proclist.do_kmc_steps(5000)
    o.do_steps(5000)
    assert base.get_kmc_step() == o.kmc_step == 5000
    assert abs(base.get_kmc_time() - o.kmc_time) <= 1e-12 * o.kmc_time
    assert [base.get_procstat(i + 1) for i in range(P)] == list(o.procstat)
    assert [base.get_nrofsites(i + 1) for i in range(P)] == list(o.nr_of_sites)
    np.testing.assert_allclose([base.get_integ_rate(i + 1) for i in range(P)], o.integ_rates, rtol=1e-10)
    np.testing.assert_allclose(proclist.get_occupation(), o.occupation, atol=1e-15)
    av = o.avail_sites
    assert all(base.get_avail_site(p, k, 1) == av[p - 1, k - 1, 0] for p in (1, 3, P) for k in (1, 2, 17))
    assert [lattice.get_species(lattice.calculate_nr2lattice(n)) for n in (1, 7, 400)] == [o.lattice[n - 1] for n in (1, 7, 400)]

    # KMC_Model.put: get_species / replace_species from Python, then _adjust_database's full touch-up pass
    site = [3, 4, 0, 1]
    old = lattice.get_species(site)
    new = (old + 1) % len(ir["species"])
    lattice.replace_species(site, old, new)
    with pytest.raises(RuntimeError):
        proclist.do_kmc_steps(1)               # edited lattice without the touch-up pass: refused
    for x in range(20):
        for y in range(20):
            getattr(proclist, "touchup_" + ir["sites"][0].lower())([x, y, 0, 1])
    base.update_accum_rate()
    spec = o.lattice.copy()
    spec[lattice.calculate_lattice2nr(site) - 1] = new
    assert o.set_configuration(spec) == 0
    assert lattice.get_species(site) == new
    proclist.do_kmc_steps(1500)
    o.do_steps(1500)
    assert [base.get_procstat(i + 1) for i in range(P)] == list(o.procstat)
    assert np.array_equal(np.asarray([lattice.get_species(lattice.calculate_nr2lattice(n)) for n in range(1, 401)]), o.lattice)
    # base.set_kmc_time (time-overrun reset of get_atoms)
    base.set_kmc_time(0.0)
    assert base.get_kmc_time() == 0.0


@pytest.mark.parametrize("kmc_model", ["pairwise_otf_otf"], indirect=True)
def test_otf_proclist_pars_on_cuda_match_the_oracle(kmc_model):
    """What kmos.run.set_rate_constants does for an otf model (kmos/run/__init__.py:2383-2437): set_rate_const
    per process, update_user_parameter per parameter, recalculate_rates_matrix -- then byst_/rate_<proc> of
    KMC_Model.rate_constants, a mid-run parameter change, and the trajectory against the oracle."""
    from kmos_b200 import otf, tables
    from kmc_model import proclist_pars
    base, lattice, proclist = kmc_model.base, kmc_model.lattice, kmc_model.proclist
    ir, blob, info = load_model("pairwise_otf_otf")
    P = len(ir["procs"])
    r = np.asarray(rates_mod.model_rates(ir))
    proclist.init([12, 10], "kmc_model", lattice.default_layer, 7, True)
    for i in range(P):
        base.set_rate_const(i + 1, float(r[i]))
    for k, name in enumerate(ir["userpar"]):
        assert getattr(proclist_pars, name.lower()) == k + 1
        proclist_pars.update_user_parameter(k + 1, otf.user_parameters(ir)[0][k])
        assert proclist_pars.get_user_parameter(k + 1) == otf.user_parameters(ir)[0][k]
    assert ir["chempots"] == ["mu_COgas"] and proclist_pars.mu_cogas == 1
    proclist_pars.update_chempot(1, -0.5)                     # chempots(1): declared, used by no rate_<proc>
    proclist.recalculate_rates_matrix()
    lut = otf.build_lut(ir, info, r)
    o = oracle.Oracle(blob, [12, 10], seed=7, replica=0, rates=r, lut=lut)

    assert "nr_CO_1nn" in "".join(proclist_pars.byst_co_desorption).split()
    p_des = proclist.co_desorption
    g = info["gr"]["gr_co_desorption"] if "gr_co_desorption" in info["gr"] else \
        [v for k, v in info["gr"].items() if k.lower() == "gr_co_desorption"][0]
    for k in range(5):
        assert proclist_pars.rate_co_desorption(np.array([k])) == lut[g["lut_offset"] + k]

    def check(steps):
        proclist.do_kmc_steps(steps)
        o.do_steps(steps)
        assert base.get_kmc_step() == o.kmc_step
        assert abs(base.get_kmc_time() - o.kmc_time) <= 1e-12 * o.kmc_time
        assert [base.get_procstat(i + 1) for i in range(P)] == list(o.procstat)
        assert [base.get_nrofsites(i + 1) for i in range(P)] == list(o.nr_of_sites)
        n = 12 * 10 * lattice.spuck
        got = [lattice.get_species(lattice.calculate_nr2lattice(k)) for k in range(1, n + 1)]
        assert np.array_equal(np.asarray(got), o.lattice)

    check(3000)
    # KMC_Model.parameters.<name> = value: update_user_parameter + recalculate_rates_matrix with new tables
    k = [n.lower() for n in ir["userpar"]].index("e_co_nn")
    proclist_pars.update_user_parameter(k + 1, 0.05)
    proclist.recalculate_rates_matrix()
    lut2 = otf.build_lut(ir, info, r, {ir["userpar"][k]: 0.05})
    assert not np.array_equal(lut, lut2)
    o.set_lut(lut2)
    o.recalculate_rates_matrix()
    assert proclist_pars.rate_co_desorption(np.array([2])) == lut2[g["lut_offset"] + 2]
    check(3000)
    base.update_accum_rate()
    o.L.kmos_oracle_update_accum_rate(o.h)
    assert abs(base.get_accum_rate(p_des) - o.accum_rates[p_des - 1]) <= 1e-12 * abs(o.accum_rates[p_des - 1])
