"""Test double for kmos_b200.engine.Batch on top of the CPU oracle (test infrastructure only).

tests/test_dropin_reference_run.py drives the *unmodified* reference front-end (kmos.run.KMC_Model) through the
kmc_model drop-in package.  That needs the reference checkout, which exists in the authoring container but not
on the GPU box -- and the container has no GPU.  So the container test plugs this double in through
kmc_model._runtime.batch_factory to check the drop-in's *API surface* against the real kmos.run, and
tests/test_gpu_dropin.py checks the same package on CUDA against the oracle.  The product never loads this."""
import numpy as np

from kmos_b200 import tables
from oracle import oracle


class OracleBatch(object):
    R = 1

    def __init__(self, ir, size, seed, layer, rng=oracle.RNG_PHILOX):
        self.ir = ir
        self.blob, self.info = tables.build_blob(ir)
        self.P = len(ir["procs"])
        self._rates = np.zeros(self.P)
        self.o = oracle.Oracle(self.blob, size, seed=seed, replica=0, rng=rng, rates=self._rates, layer=layer)
        self.volume = self.o.volume

    def close(self):
        self.o = None

    def do_steps(self, n):
        self.o.do_steps(int(n))

    def get_next_kmc_step(self):
        p, s, _st = self.o.get_next_kmc_step()
        return np.array([p], np.int32), np.array([s], np.int32)

    def run_proc_nr(self, proc, site):
        self.o.run_proc_nr(int(proc), int(site))

    def set_rate_const(self, proc, rate, replica=0):
        self._rates[int(proc) - 1] = float(rate)
        self.o.set_rates(self._rates)

    def set_otf_lut(self, lut):
        """kmos_b200_set_otf_lut: new rate table + recalculate_rates_matrix (include/kmos_b200.h)."""
        self.o.set_lut(np.asarray(lut, dtype=np.float64).reshape(-1))
        self.o.recalculate_rates_matrix()

    def set_kmc_time(self, t):
        self.o.set_kmc_time(float(np.asarray(t).reshape(-1)[0]))

    def set_configuration(self, species, replica=0, layer=None):
        assert self.o.set_configuration(np.asarray(species, dtype=np.int32)) == 0

    def avail_sites(self, replica):
        return self.o.avail_sites

    kmc_time = property(lambda self: np.array([self.o.kmc_time]))
    kmc_time_step = property(lambda self: np.array([self.o.kmc_time_step]))
    kmc_step = property(lambda self: np.array([self.o.kmc_step]))
    procstat = property(lambda self: self.o.procstat[None, :])
    integ_rates = property(lambda self: self.o.integ_rates[None, :])
    nr_of_sites = property(lambda self: self.o.nr_of_sites[None, :])
    rates = property(lambda self: self._rates[None, :].copy())
    lattice = property(lambda self: self.o.lattice[None, :])
    occupation = property(lambda self: self.o.occupation[None, :, :])

    @property
    def accum_rates(self):
        self.o.L.kmos_oracle_update_accum_rate(self.o.h)
        return self.o.accum_rates[None, :]
