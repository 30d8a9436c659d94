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
    return sorted(["update_user_parameter", "update_chempot"] + [n.lower() for n in rt.RT.ir.get("userpar", [])])


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


def update_chempot(index, value):
    raise NotImplementedError("chemical potentials are tabulated from T and p by kmos_b200.otf")


def __getattr__(name):
    names = [n.lower() for n in rt.RT.ir.get("userpar", [])]
    if name.lower() in names:
        return names.index(name.lower()) + 1
    raise AttributeError(name)
