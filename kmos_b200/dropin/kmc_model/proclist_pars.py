"""``kmc_model.proclist_pars`` (otf only): user parameters feed the host-tabulated rate tables
(kmos_b200/otf.py); a changed parameter re-tabulates and re-sends them (kmos_b200_set_otf_lut, which also
refreshes every registered rate: proclist.recalculate_rates_matrix)."""
import numpy as np

from . import _runtime as rt

if rt.RT.ir is None:
    rt.load()
if rt.RT.ir["backend"] != "otf":  # the generated module exists for otf models only (kmos/run/__init__.py:110-113)
    raise ImportError("proclist_pars: not an otf model")

_overrides = {}


def __dir__():
    procs = [p.lower() for p in rt.RT.ir["procs"]]
    return sorted(["update_user_parameter", "get_user_parameter"] +
                  (["update_chempot"] if rt.RT.ir.get("chempots") else []) +
                  [n.lower() for n in rt.RT.ir.get("userpar", []) + rt.RT.ir.get("chempots", [])] +
                  ["byst_" + p for p in procs] + ["rate_" + p for p in procs])


def _send():
    from kmos_b200 import otf, tables
    ir = rt.RT.ir
    _blob, info = tables.build_blob(ir)
    rates = rt.batch().rates[0]
    rt.batch().set_otf_lut(np.asarray(otf.build_lut(ir, info, rates, _overrides))[None, :])


def update_user_parameter(index, value):
    names = rt.RT.ir.get("userpar", [])
    _overrides[names[int(index) - 1]] = float(value)
    _send()


def get_user_parameter(index):
    """val = userpar(param) (kmos/io/__init__.py:2766-2774)."""
    from kmos_b200 import otf
    return float(otf.user_parameters(rt.RT.ir, _overrides)[0][int(index) - 1])


def _update_chempot(index, value):
    """chempots(index) = val: the value kmos.run.set_rate_constants evaluated replaces the tabulated one."""
    _overrides[rt.RT.ir["chempots"][int(index) - 1]] = float(value)
    _send()


def __getattr__(name):
    # the generated module has update_chempot only when the model uses chemical potentials
    # (kmos/io/__init__.py:2776-2786); kmos.run probes for it with hasattr
    if name == "update_chempot" and rt.RT.ir.get("chempots"):
        return _update_chempot
    low = name.lower()
    if low.startswith("byst_") or low.startswith("rate_"):
        # byst_<proc>: names of the bystander counters; rate_<proc>(nr_vars): the rate for a given environment
        by_lower = {p.lower(): p for p in rt.RT.ir["procs"]}
        proc = by_lower.get(low[5:])
        if proc is None:
            raise AttributeError(name)
        if low.startswith("byst_"):
            return {k.lower(): v for k, v in rt.RT.ir.get("byst", {}).items()}.get(proc.lower(), "")
        from kmos_b200 import otf
        return otf.rate_function(rt.RT.ir, proc, rt.batch().rates[0], _overrides)
    for names in (rt.RT.ir.get("userpar", []), rt.RT.ir.get("chempots", [])):
        low = [n.lower() for n in names]
        if name.lower() in low:
            return low.index(name.lower()) + 1
    raise AttributeError(name)
