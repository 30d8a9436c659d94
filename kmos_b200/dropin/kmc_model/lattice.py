"""``kmc_model.lattice``: the f2py view of kmos/fortran_src/lattice.mpy that kmos.run uses."""
import numpy as np

from . import _runtime as rt


def get_species(site):
    return int(rt.lattice_host()[rt.lattice2nr(site) - 1])


def replace_species(site, old_species, new_species):
    rt.stage_species(site, old_species, new_species)


def calculate_lattice2nr(site):
    return rt.lattice2nr(site)


def calculate_nr2lattice(nr):
    return rt.nr2lattice(nr)


def deallocate_system():
    rt.deallocate()


def __dir__():
    ir = rt.RT.ir
    return sorted(["get_species", "replace_species", "calculate_lattice2nr", "calculate_nr2lattice",
                   "deallocate_system", "system_size", "spuck", "model_dimension", "default_layer", "nr_of_layers",
                   "substrate_layer", "unit_cell_size", "site_positions"]
                  + [x.lower() for x in ir["layers"]] + [s.lower() for s in ir["sites"]])


def __getattr__(name):
    ir = rt.RT.ir
    if name == "system_size":
        return np.array(rt.RT.size if rt.RT.size is not None else [0, 0, 0])
    if name in ("spuck", "model_dimension", "default_layer"):
        return ir[name]
    if name == "nr_of_layers":
        return len(ir["layers"])
    if name == "substrate_layer":
        return ir.get("substrate_layer", ir["default_layer"])
    if name == "unit_cell_size":
        return np.array(ir.get("unit_cell_size", np.eye(3).tolist()), dtype=float)
    if name == "site_positions":
        return np.array(ir.get("site_positions", np.zeros((ir["spuck"], 3)).tolist()), dtype=float)
    low = name.lower()
    for i, layer in enumerate(ir["layers"]):       # layer constants: lattice.<layer> = index
        if layer.lower() == low:
            return i
    for i, site in enumerate(ir["sites"]):         # site constants: lattice.<layer>_<site> = 1-based site type
        if site.lower() == low:
            return i + 1
    raise AttributeError(name)
