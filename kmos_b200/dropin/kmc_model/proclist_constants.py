"""``kmc_model.proclist_constants``: process numbers (1-based) and species ids (0-based) by lower-case name,
nr_of_proc, nr_of_species (what the generated proclist_constants.f90 declares)."""
from . import _runtime as rt


def _table():
    ir = rt.RT.ir
    out = {"nr_of_proc": len(ir["procs"]), "nr_of_species": len(ir["species"]),
           "default_species": ir.get("default_species", 0)}
    for i, sp in enumerate(ir["species"]):
        out[sp.lower()] = i
    for i, p in enumerate(ir["procs"]):
        out[p.lower()] = i + 1
    return out


def __getattr__(name):
    t = _table()
    if name.lower() in t:
        return t[name.lower()]
    raise AttributeError(name)


def __dir__():
    return sorted(_table())
