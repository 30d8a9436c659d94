"""``kmc_model.base``: the f2py view of kmos/fortran_src/base.mpy that kmos.run uses."""
import numpy as np

from . import _runtime as rt


def is_allocated():
    return rt.RT.batch is not None


def get_null_species():
    return -1 if rt.RT.ir.get("null_species", -1) is None else int(rt.RT.ir.get("null_species", -1))


def get_volume():
    return rt.volume()


def get_kmc_time():
    rt.clean()
    return float(rt.batch().kmc_time[0])


def get_kmc_time_step():
    rt.clean()
    return float(rt.batch().kmc_time_step[0])


def get_kmc_step():
    rt.clean()
    return int(rt.batch().kmc_step[0])


def set_kmc_time(t):
    rt.clean()
    rt.batch().set_kmc_time(np.array([float(t)]))


def get_procstat(proc):
    rt.clean()
    return int(rt.batch().procstat[0, int(proc) - 1])


def get_integ_rate(proc):
    rt.clean()
    return float(rt.batch().integ_rates[0, int(proc) - 1])


def get_nrofsites(proc):
    rt.clean()
    return int(rt.batch().nr_of_sites[0, int(proc) - 1])


def get_rate(proc):
    return float(rt.batch().rates[0, int(proc) - 1])


def get_accum_rate(proc):
    rt.clean()
    return float(rt.batch().accum_rates[0, int(proc) - 1])


def set_rate_const(proc, rate):
    rt.batch().set_rate_const(int(proc), float(rate), replica=0)


def get_avail_site(proc, field, switch):
    rt.clean()
    return int(rt.batch().avail_sites(0)[int(proc) - 1, int(field) - 1, int(switch) - 1])


def update_accum_rate():
    """The device recomputes accum_rates whenever it is read or stepped; what remains of the reference's call
    is its role as the end marker of KMC_Model._adjust_database."""
    rt.flush()


def update_integ_rate():
    """base.update_integ_rate at reset time adds accum * kmc_time_step with kmc_time_step = 0: nothing to do;
    while stepping the device does it every step (kmos/fortran_src/base.mpy:626-645)."""
    return None


def __getattr__(name):  # f2py exposes module variables as attributes
    if name == "null_species":
        return get_null_species()
    raise AttributeError(name)
