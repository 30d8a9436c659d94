"""``kmc_model.proclist``: the f2py view of the generated proclist module that kmos.run uses
(kmos/fortran_src/proclist_generic_subroutines.mpy, kmos/io/__init__.py:305-465)."""
import numpy as np

from . import _runtime as rt
from . import proclist_constants as _constants

seed = np.array(1)


def init(size, system_name="kmc_model", layer=0, seed_in=1, no_banner=True):
    """allocate_system + initialize_state; the banner is the reference's business, not ours."""
    global seed
    seed = np.array(int(seed_in))
    rt.allocate(np.asarray(size).reshape(-1), int(layer), int(seed_in))


def do_kmc_steps(n):
    rt.clean()
    rt.batch().do_steps(int(n))


def do_kmc_step():
    do_kmc_steps(1)


def get_next_kmc_step():
    rt.clean()
    proc, site = rt.batch().get_next_kmc_step()
    return np.int32(proc[0]), np.int32(site[0])


def run_proc_nr(proc, site):
    rt.clean()
    rt.batch().run_proc_nr(int(proc), int(site))


def get_occupation():
    """occupation[n_species][spuck] (proclist_generic_subroutines.mpy:113-158)."""
    rt.clean()
    return np.asarray(rt.batch().occupation[0])


def touchup_cell(cell):
    rt.note_touchup(cell)


def recalculate_rates_matrix():
    """otf: re-tabulate gr_<proc> from the current rate constants and user parameters and send the table;
    kmos_b200_set_otf_lut then refreshes every registered rate and re-adds the rows (include/kmos_b200.h)."""
    if rt.RT.ir["backend"] == "otf":
        from . import proclist_pars
        proclist_pars._send()


def __dir__():
    names = ["init", "do_kmc_steps", "do_kmc_step", "get_next_kmc_step", "run_proc_nr", "get_occupation", "seed",
             "nr_of_proc", "backend", "touchup_cell"]
    names += ["touchup_" + s.lower() for s in rt.RT.ir["sites"]]
    if rt.RT.ir["backend"] == "otf":
        names.append("recalculate_rates_matrix")
    return sorted(set(names) | set(dir(_constants)))


def __getattr__(name):
    low = name.lower()
    if low == "nr_of_proc":
        return len(rt.RT.ir["procs"])
    if low == "backend":
        return rt.RT.ir["backend"]
    if low.startswith("touchup_"):  # touchup_<layer>_<site>(site): one call per site in _adjust_database
        return lambda site: rt.note_touchup(site)
    try:
        return getattr(_constants, name)
    except AttributeError:
        raise AttributeError("%s not found in kmc_model.proclist" % name)
