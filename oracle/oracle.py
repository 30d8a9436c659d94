"""ctypes front-end of the CPU oracle (oracle/kmos_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkmos_oracle.so")

RNG_PHILOX, RNG_GFORTRAN = 0, 1
OK, DEADLOCK, SPECIES_MISMATCH, CAPACITY, BAD_MODEL = range(5)


def build(force=False):
    src = os.path.join(HERE, "kmos_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        i32p, i64p, f64p = (np.ctypeslib.ndpointer(dtype=d, flags="C_CONTIGUOUS")
                            for d in (np.int32, np.int64, np.float64))
        L.kmos_oracle_create.restype = C.c_void_p
        L.kmos_oracle_create.argtypes = [i32p, C.c_int64, i32p]
        L.kmos_oracle_destroy.argtypes = [C.c_void_p]
        L.kmos_oracle_seed.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32]
        L.kmos_oracle_set_rates.argtypes = [C.c_void_p, f64p]
        L.kmos_oracle_set_lut.argtypes = [C.c_void_p, f64p]
        L.kmos_oracle_recalculate_rates_matrix.argtypes = [C.c_void_p]
        L.kmos_oracle_recalculate_rates_matrix.restype = None
        L.kmos_oracle_init_state.argtypes = [C.c_void_p, C.c_int]
        L.kmos_oracle_set_configuration.argtypes = [C.c_void_p, i32p, C.c_int]
        L.kmos_oracle_do_steps.argtypes = [C.c_void_p, C.c_int64]
        L.kmos_oracle_get_next_kmc_step.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.kmos_oracle_run_proc_nr.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.kmos_oracle_update_accum_rate.argtypes = [C.c_void_p]
        L.kmos_oracle_interval_search_real.argtypes = [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_int32, C.c_double]
        L.kmos_oracle_interval_search_real.restype = C.c_int
        L.kmos_oracle_volume.argtypes = [C.c_void_p]
        L.kmos_oracle_nproc.argtypes = [C.c_void_p]
        L.kmos_oracle_status.argtypes = [C.c_void_p, i32p]
        for name in ("kmc_time", "kmc_time_step"):
            f = getattr(L, "kmos_oracle_" + name)
            f.restype = C.c_double
            f.argtypes = [C.c_void_p]
        L.kmos_oracle_kmc_step.restype = C.c_int64
        L.kmos_oracle_kmc_step.argtypes = [C.c_void_p]
        L.kmos_oracle_get_lattice.argtypes = [C.c_void_p, i32p]
        L.kmos_oracle_get_procstat.argtypes = [C.c_void_p, i64p]
        L.kmos_oracle_get_nr_of_sites.argtypes = [C.c_void_p, i32p]
        L.kmos_oracle_get_integ_rates.argtypes = [C.c_void_p, f64p]
        L.kmos_oracle_get_accum_rates.argtypes = [C.c_void_p, f64p]
        L.kmos_oracle_get_avail_sites.argtypes = [C.c_void_p, i32p]
        L.kmos_oracle_get_rates_matrix_row.argtypes = [C.c_void_p, C.c_int, f64p]
        L.kmos_oracle_get_occupation.argtypes = [C.c_void_p, f64p]
        L.kmos_oracle_get_counters.argtypes = [C.c_void_p, i64p]
        L.kmos_oracle_reset_counters.argtypes = [C.c_void_p]
        L.kmos_oracle_philox_step.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, f64p]
        _lib = L
    return _lib


def interval_search_real(arr, value):
    """base.interval_search_real (base.mpy:1234-1338) on a float64 array: 1-based index, 0 where the reference stops."""
    a = np.ascontiguousarray(arr, dtype=np.float64)
    return int(lib().kmos_oracle_interval_search_real(a, a.size, float(value)))


def philox_step(seed, replica, step):
    out = np.zeros(3)
    lib().kmos_oracle_philox_step(seed, replica, step, out)
    return out


class Oracle(object):
    """One replica of the reference semantics on the CPU."""

    def __init__(self, blob, size, seed=1, replica=0, rng=RNG_PHILOX, rates=None, lut=None, layer=None,
                 init=True):
        self.L = lib()
        self.blob = np.ascontiguousarray(blob, dtype=np.int32)
        size3 = np.ones(3, dtype=np.int32)
        size = np.atleast_1d(np.asarray(size, dtype=np.int32))
        size3[:len(size)] = size
        self.h = self.L.kmos_oracle_create(self.blob, self.blob.size, size3)
        if not self.h:
            raise ValueError("bad model blob")
        self.size = size3
        self.n_proc = self.L.kmos_oracle_nproc(self.h)
        self.volume = self.L.kmos_oracle_volume(self.h)
        self.n_species = int(self.blob[3])
        self.spuck = int(self.blob[5])
        self.default_layer = int(self.blob[9])
        self.layer = self.default_layer if layer is None else layer
        self.L.kmos_oracle_seed(self.h, rng, seed, replica)
        if rates is not None:
            self.set_rates(rates)
        if lut is not None:
            self.set_lut(lut)
        if init:
            st = self.L.kmos_oracle_init_state(self.h, self.layer)
            if st != OK:
                raise RuntimeError("init_state status %d" % st)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.kmos_oracle_destroy(self.h)
            self.h = None

    def set_rates(self, rates):
        r = np.ascontiguousarray(rates, dtype=np.float64)
        assert r.size == self.n_proc
        self.L.kmos_oracle_set_rates(self.h, r)

    def set_lut(self, lut):
        self.L.kmos_oracle_set_lut(self.h, np.ascontiguousarray(lut, dtype=np.float64))

    def recalculate_rates_matrix(self):
        """proclist.recalculate_rates_matrix: what KMC_Model.set_rate_constants ends with for otf models."""
        self.L.kmos_oracle_recalculate_rates_matrix(self.h)

    def set_configuration(self, species):
        s = np.ascontiguousarray(species, dtype=np.int32)
        assert s.size == self.volume
        return self.L.kmos_oracle_set_configuration(self.h, s, self.layer)

    def do_steps(self, n):
        return self.L.kmos_oracle_do_steps(self.h, int(n))

    def get_next_kmc_step(self):
        p, s = C.c_int32(0), C.c_int32(0)
        st = self.L.kmos_oracle_get_next_kmc_step(self.h, C.byref(p), C.byref(s))
        return p.value, s.value, st

    def run_proc_nr(self, proc, site):
        return self.L.kmos_oracle_run_proc_nr(self.h, proc, site)

    @property
    def status(self):
        err = np.zeros(5, dtype=np.int32)
        return self.L.kmos_oracle_status(self.h, err), err

    @property
    def kmc_time(self):
        return self.L.kmos_oracle_kmc_time(self.h)

    @property
    def kmc_time_step(self):
        return self.L.kmos_oracle_kmc_time_step(self.h)

    def set_kmc_time(self, t):
        self.L.kmos_oracle_set_kmc_time(self.h, float(t))

    @property
    def kmc_step(self):
        return self.L.kmos_oracle_kmc_step(self.h)

    def _get(self, fn, shape, dtype):
        out = np.zeros(shape, dtype=dtype)
        getattr(self.L, "kmos_oracle_get_" + fn)(self.h, out)
        return out

    lattice = property(lambda self: self._get("lattice", self.volume, np.int32))
    procstat = property(lambda self: self._get("procstat", self.n_proc, np.int64))
    nr_of_sites = property(lambda self: self._get("nr_of_sites", self.n_proc, np.int32))
    integ_rates = property(lambda self: self._get("integ_rates", self.n_proc, np.float64))
    accum_rates = property(lambda self: self._get("accum_rates", self.n_proc, np.float64))
    avail_sites = property(lambda self: self._get("avail_sites", (self.n_proc, self.volume, 2), np.int32))
    occupation = property(lambda self: self._get("occupation", (self.n_species, self.spuck), np.float64))

    def rates_matrix_row(self, proc):
        """otf: rates_matrix(proc, 1..volume+1) (last entry: the row total)."""
        out = np.zeros(self.volume + 1)
        self.L.kmos_oracle_get_rates_matrix_row(self.h, int(proc), out)
        return out
    counters = property(lambda self: dict(zip(("n_rs", "n_chk", "n_del", "n_gs", "n_add", "n_upd"),
                                              self._get("counters", 6, np.int64).tolist())))

    def reset_counters(self):
        self.L.kmos_oracle_reset_counters(self.h)
