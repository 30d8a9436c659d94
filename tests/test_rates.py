"""Host-side rate preparation: kmos_b200.rates.evaluate_rate_expression mirrors kmos.evaluate_rate_expression
(kmos/__init__.py:67-189).  The vectors of the reference's own tests/test_evaluate_rate_expression.py are restated
here; where the reference checkout is importable (this container, not the GPU box) the two implementations are
also run side by side on every rate expression of the RuO2 / ZGB / pairwise fixtures that needs no JANAF data."""
import math
import os
import sys

import pytest

from conftest import REPO, load_model
from kmos_b200 import rates

REF = "/root/reference"


def test_reference_unit_vectors():
    ev = rates.evaluate_rate_expression
    assert ev("1.5e-3") == pytest.approx(1.5e-3)
    assert ev("2 * 3 + 4") == pytest.approx(10.0)
    assert ev("exp(1)") == pytest.approx(math.e, rel=1e-12)
    assert ev("T * 2", {"T": {"value": 600}, "p_CO": {"value": 1.0}}) == pytest.approx(1200.0)
    assert ev("kboltzmann * 600") == pytest.approx(1.3806488e-23 * 600, rel=1e-6)
    r = ev("1/(beta*h)*exp(-beta*0.9*eV)", {"T": {"value": 600}})
    assert 0 < r < 1e20
    assert ev("") == 0.0


def test_model_rates_are_finite_and_positive():
    for name in ("ruo2_local_smart", "zgb_local_smart", "pairwise_lat_int", "ab_local_smart"):
        ir, _blob, _info = load_model(name)
        r = rates.model_rates(ir)
        assert len(r) == len(ir["procs"]) and all(math.isfinite(x) and x >= 0 for x in r)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "kmos")), reason="reference checkout absent")
def test_side_by_side_with_the_reference_implementation():
    sys.path.insert(0, os.path.join(REPO, "tools", "ase_shim"))
    sys.path.insert(0, REF)
    try:
        import kmos
    finally:
        sys.path.remove(REF)
    checked = 0
    for name in ("ruo2_local_smart", "zgb_local_smart", "pairwise_lat_int", "ab_local_smart"):
        ir, _blob, _info = load_model(name)
        params = {k: {"value": v["value"]} for k, v in ir["parameters"].items()}
        for p in ir["process_defs"]:
            expr = p["rate_constant"]
            if not expr or "mu_" in expr:  # mu_* needs the JANAF tables, which are not vendored
                continue
            ours = rates.evaluate_rate_expression(expr, params)
            ref = kmos.evaluate_rate_expression(rate_expr=expr, parameters=params)
            assert ours == pytest.approx(ref, rel=1e-13), (name, p["name"], expr)
            checked += 1
    assert checked >= 40


def test_mu_tokens_use_a_provider_or_warn_about_the_standin():
    """a drop-in KMC_Model must not silently replace the JANAF chemical potentials.  `mu` accepts a
    kmos.species-compatible provider; without one (and without an importable reference) the closed-form stand-in
    is used *with a warning*; a gas without a table evaluates to 0 with a warning, as in the reference
    (kmos/__init__.py:128-160)."""
    import warnings

    import pytest

    params = {"T": {"value": 500.0}, "p_COgas": {"value": 2.0}, "p_Xegas": {"value": 1.0}}

    class Gas(object):
        def mu(self, T, p):
            return -1.25 + 1e-3 * T * p

    class Provider(object):
        COgas = Gas()

    got = rates.evaluate_rate_expression("exp(mu_COgas)", params, mu=Provider())
    assert got == pytest.approx(math.exp(-1.25 + 1e-3 * 500.0 * 2.0), rel=1e-15)
    with pytest.warns(rates.MuStandinWarning, match="No JANAF table"):
        assert rates.evaluate_rate_expression("1 + mu_Xegas", params, mu=Provider()) == 1.0
    if rates._reference_species() is None:  # the GPU box and this container: no kmos installation
        with pytest.warns(rates.MuStandinWarning, match="stand-in"):
            v = rates.evaluate_rate_expression("mu_COgas", params)
        assert v == rates.standin_mu("COgas", 500.0, 2.0)
    with warnings.catch_warnings():
        warnings.simplefilter("error")  # the explicit stand-in (synthetic workloads) stays silent
        rates.evaluate_rate_expression("mu_COgas", params, mu=rates.standin_mu)
    with pytest.raises(KeyError):
        rates.evaluate_rate_expression("mu_O2gas", params, mu=rates.standin_mu)


def test_vectorised_rate_preparation_matches_the_point_by_point_one():
    """workloads.ruo2_grid / otf LUTs over a whole sweep in one pass per expression."""
    import time

    import numpy as np

    from kmos_b200 import otf as otf_mod, tables, workloads
    ir, _blob, _info = load_model("ruo2_local_smart")
    Ts, ps = np.linspace(450.0, 650.0, 5), np.logspace(-2, 2, 4)
    TT, PP = np.meshgrid(Ts, ps, indexing="ij")
    grid = rates.model_rates_grid(ir, {"T": TT.reshape(-1), "p_COgas": PP.reshape(-1)},
                                  overrides={"p_O2gas": 1.0}, mu=rates.standin_mu)
    ref = np.asarray([rates.model_rates(ir, {"T": float(T), "p_COgas": float(p), "p_O2gas": 1.0}, mu=rates.standin_mu)
                      for T in Ts for p in ps])
    assert grid.shape == ref.shape == (20, 36)
    np.testing.assert_allclose(grid, ref, rtol=1e-13, atol=0)
    t0 = time.time()
    r, group_of, _d = workloads.ruo2_grid(ir)
    assert r.shape == (16384, 36) and group_of[-1] == 255 and time.time() - t0 < 5.0
    # otf tables for many replicas at once
    ir2, blob2, info2 = load_model("intzgb_otf")
    rr = np.exp(np.random.RandomState(0).uniform(-1, 1, (7, len(ir2["procs"]))))
    batch = otf_mod.build_lut_batch(ir2, info2, rr)
    one = np.stack([otf_mod.build_lut(ir2, info2, rr[i]) for i in range(7)])
    np.testing.assert_allclose(batch, one, rtol=1e-14, atol=0)
