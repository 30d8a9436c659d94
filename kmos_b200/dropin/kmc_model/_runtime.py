"""Shared state of the kmc_model drop-in: the model tables and the one-replica batch behind the f2py-shaped
functions (the f2py module is a process-wide singleton as well: kmos/run/__init__.py:147-150)."""
import json
import os
import sys

import numpy as np


class Runtime(object):
    def __init__(self):
        self.ir = None
        self.tables_path = None
        self.model = None
        self.batch = None
        self.size = None
        self.layer = 0
        self.seed = 1
        self.host_lattice = None  # species per site while lattice.replace_species edits are pending
        self.dirty = False
        self.touched = []


RT = Runtime()
# tests replace this to drive the package without a GPU (tests/oracle_engine.py); the product path is CUDA only
batch_factory = None


def find_tables():
    p = os.environ.get("KMOS_B200_MODEL")
    if p:
        return p
    for d in [os.getcwd()] + list(sys.path):
        d = d or os.getcwd()
        if os.path.isfile(os.path.join(d, "kmc_settings.py")) or d == os.getcwd():
            cand = os.path.join(d, "model_tables.json")
            if os.path.isfile(cand):
                return cand
    raise ImportError("kmc_model (kmos_b200 drop-in): no model tables; set KMOS_B200_MODEL or put the "
                      "model_tables.json written by kmos_b200.export next to kmc_settings.py")


def load(path=None):
    from kmos_b200 import tables
    RT.tables_path = path or find_tables()
    RT.ir = tables.load_ir(RT.tables_path)
    RT.model = None
    RT.batch = None
    return RT.ir


def allocate(size, layer, seed):
    """proclist.init: allocate_system + initialize_state (proclist_generic_subroutines.mpy:160-304)."""
    from kmos_b200 import engine
    deallocate()
    dim = RT.ir["model_dimension"]
    RT.size = np.ones(3, dtype=np.int64)
    RT.size[:dim] = np.asarray(size, dtype=np.int64).reshape(-1)[:dim]
    RT.layer, RT.seed = int(layer), int(seed)
    seeds = np.array([RT.seed], dtype=np.uint64)
    if batch_factory is not None:
        RT.model = None
        RT.batch = batch_factory(RT.ir, RT.size[:dim].astype(np.int32), RT.seed, RT.layer)
    else:
        RT.model = engine.Model(ir=RT.ir)
        RT.batch = engine.Batch(RT.model, 1, RT.size[:dim].astype(np.int32), seeds=seeds, layer=RT.layer)
    RT.host_lattice, RT.dirty, RT.touched = None, False, []


def deallocate():
    if RT.batch is not None:
        RT.batch.close()
    if RT.model is not None:
        RT.model.close()
    RT.batch = RT.model = None
    RT.host_lattice, RT.dirty, RT.touched = None, False, []


def batch():
    if RT.batch is None:
        raise RuntimeError("kmc_model: the system is not allocated (proclist.init has not been called)")
    return RT.batch


def ncells():
    return int(RT.size[0] * RT.size[1] * RT.size[2])


def volume():
    return ncells() * RT.ir["spuck"]


def lattice2nr(site):
    x, y, z, n = (int(v) for v in site)
    Lx, Ly, Lz = (int(v) for v in RT.size)
    return ((x % Lx) + Lx * ((y % Ly) + Ly * (z % Lz))) * RT.ir["spuck"] + n


def nr2lattice(nr):
    cell, n = divmod(int(nr) - 1, RT.ir["spuck"])
    Lx, Ly = int(RT.size[0]), int(RT.size[1])
    return np.array([cell % Lx, (cell // Lx) % Ly, cell // (Lx * Ly), n + 1])


def lattice_host():
    """Species per site (site-number order), fetched once and kept while edits are pending."""
    if RT.host_lattice is None:
        RT.host_lattice = np.array(batch().lattice[0], dtype=np.int32)
    return RT.host_lattice


def stage_species(site, old_species, new_species):
    """lattice.replace_species from Python (KMC_Model._put / _set_configuration): staged on the host until the
    touch-up pass that has to follow (kmos/run/__init__.py:1439-1457) sends it to the device in one call."""
    lat = lattice_host()
    i = lattice2nr(site) - 1
    if int(lat[i]) != int(old_species):
        raise RuntimeError("replace_species: site %s holds species %d, not %d" % (list(site), lat[i], old_species))
    lat[i] = int(new_species)
    RT.dirty = True


def note_touchup(cell):
    RT.touched.append(tuple(int(v) for v in cell[:3]))


def flush():
    """End of KMC_Model._adjust_database (its base.update_accum_rate call): if the Python side has edited the
    lattice and touched up every cell, x outermost, hand the configuration to the device, which replays exactly
    that pass (kmos_b200_set_configuration = lattice + _adjust_database)."""
    if not RT.dirty and not RT.touched:
        return
    Lx, Ly, Lz = (int(v) for v in RT.size)
    spuck = RT.ir["spuck"]
    per_cell = 1 if RT.ir["backend"] in ("lat_int", "otf") else spuck
    full = [(x, y, z) for x in range(Lx) for y in range(Ly) for z in range(Lz) for _ in range(per_cell)]
    if RT.touched != full:
        RT.touched = []
        raise NotImplementedError("kmc_model (kmos_b200): touchup_* must be called for every site, x outermost, "
                                  "as KMC_Model._adjust_database does; partial touch-ups are not supported")
    batch().set_configuration(lattice_host(), replica=0, layer=RT.layer)
    RT.host_lattice, RT.dirty, RT.touched = None, False, []


def clean():
    """Called by everything that steps or reads derived state: pending edits must have been flushed."""
    if RT.touched:
        flush()
    if RT.dirty:
        raise RuntimeError("kmc_model: the lattice was edited (replace_species) without the touch-up pass "
                           "KMC_Model._adjust_database performs")
    RT.host_lattice = None
