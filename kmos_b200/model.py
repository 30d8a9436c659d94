"""Host-side mirror of ``kmos.run.KMC_Model`` for the batched GPU engine.

Same names, argument meaning and output formats as the reference's Python front-end
(kmos/run/__init__.py): ``do_steps`` (:416), ``get_atoms(geometry=False)`` (:680-850, TOF and occupation
bookkeeping), ``get_std_header`` (:852), ``get_std_sampled_data`` (:865-966), ``put`` /
``_get_configuration`` / ``_set_configuration`` (:1243-1437), ``dump_config`` / ``load_config`` (:1619-1640),
``get_backend`` (:1459) and the ``base`` / ``lattice`` / ``proclist`` call surface the f2py extension offers
(kmos/run/__init__.py:86-146).  The difference is the leading replica axis: one model object steps R
independent replicas (seeds and/or parameter points); every reference-shaped method takes ``replica=``
(default 0) and has a ``*_all`` twin that returns arrays over replicas.

The geometry (ASE ``Atoms``) half of ``get_atoms`` is viewer territory and out of scope.
"""
import numpy as np

from . import capi, engine, otf as otf_mod, rates as rates_mod, tables


class _Expando(object):
    pass


class _Namespace(object):
    """f2py-module-shaped view of replica 0 (what ``from kmc_model import base, lattice, proclist`` gives)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class KMC_Model(object):
    def __init__(self, model, size=20, n_replicas=1, seeds=None, parameters=None, device=0, kernel=capi.KERNEL_AUTO,
                 random_seed=1, mu=None, cache_file=None, replica_ids=None, gpu_ids=None):
        """model: path of a rule-table JSON (what the exporter hook writes) or a parsed IR dict.
        parameters: dict of overrides, or a list of R dicts (one parameter point per replica).
        mu: chemical potentials for ``mu_<gas>`` tokens: a ``kmos.species``-compatible provider or a callable
        (gas, T, p) -> eV.  Default: the reference's kmos.species if importable, else the closed-form stand-in
        of kmos_b200.rates with a MuStandinWarning (rate constants then differ from the reference's).
        replica_ids: the replicas' indices in the Philox counter (default 0..R-1); a sweep sharded over several
        batches passes its global indices so that the shards reproduce the unsharded run.
        gpu_ids: deal the replicas to these GPUs from this one process (engine.Fleet, kmos_b200_fleet_*) instead
        of running them on `device`; every replica steps exactly as it does on a single GPU."""
        self.ir = tables.load_ir(model) if isinstance(model, str) else model
        self.model = engine.Model(ir=self.ir)
        dim = self.ir["model_dimension"]
        self.size = np.ones(3, dtype=np.int64)
        self.size[:dim] = np.broadcast_to(np.asarray(size, dtype=np.int64), (dim,))
        self.R = int(n_replicas)
        self._mu = mu
        if isinstance(parameters, (list, tuple)):
            assert len(parameters) == self.R
            self._overrides = [dict(p) for p in parameters]
        else:
            self._overrides = [dict(parameters or {}) for _ in range(self.R)]
        if seeds is None:
            seeds = np.uint64(random_seed) + np.arange(self.R, dtype=np.uint64)
        if gpu_ids is not None:
            if replica_ids is not None:
                raise ValueError("gpu_ids and replica_ids are exclusive: a fleet numbers its replicas 0..R-1")
            self.batch = engine.Fleet(self.model, self.R, self.size[:dim].astype(np.int32), gpu_ids=gpu_ids,
                                      seeds=seeds, rates=self._evaluate_rates(), lut=self._evaluate_lut(),
                                      kernel=kernel)
        else:
            self.batch = engine.Batch(self.model, self.R, self.size[:dim].astype(np.int32), device=device,
                                      seeds=seeds, replica_ids=replica_ids,
                                      rates=self._evaluate_rates(), lut=self._evaluate_lut(), kernel=kernel)
        self.species_names = list(self.ir["species"])
        self.site_names = list(self.ir["sites"])
        self.process_names = list(self.ir["procs"])
        # TOF bookkeeping (kmos/run/__init__.py:262-267)
        tof_counts = {}
        for p in self.ir["process_defs"]:
            tc = p.get("tof_count")
            if isinstance(tc, str):  # .ini models carry the dict as text: "{'TOF': 1}"
                import ast
                tc = ast.literal_eval(tc)
            if tc:
                tof_counts[p["name"].lower()] = tc
        self.tofs = sorted({name for tc in tof_counts.values() for name in tc})
        self.tof_matrix = np.zeros((len(self.tofs), self.model.n_proc))
        for i, pname in enumerate(self.process_names):
            for tof, factor in (tof_counts.get(pname.lower()) or {}).items():
                self.tof_matrix[self.tofs.index(tof), i] += factor
        self._procstat = np.zeros((self.R, self.model.n_proc), dtype=np.int64)
        self._integ = np.zeros((self.R, self.model.n_proc))
        self._time = np.zeros(self.R)
        self._steps = np.zeros(self.R, dtype=np.int64)
        self._tof_data = np.zeros((self.R, len(self.tofs)))
        self._tof_integ = np.zeros((self.R, len(self.tofs)))
        self._shim()
        if cache_file is not None:
            import os
            if os.path.exists(cache_file):
                self.load_config(cache_file)

    # ---- context manager / teardown (kmos/run/__init__.py:388-414) ----------------------------------------
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.deallocate()

    def deallocate(self):
        self.batch.close()

    # ---- parameters -> rate constants (kmos/run/__init__.py:1643-1895, 2369-2437) -------------------------
    def _evaluate_rates(self):
        return np.asarray([rates_mod.model_rates(self.ir, ov, mu=self._mu) for ov in self._overrides])

    def _evaluate_lut(self):
        if self.ir["backend"] != "otf":
            return None
        info = self.model.info
        return np.stack([otf_mod.build_lut(self.ir, info, r, ov) for r, ov in
                         zip(self._evaluate_rates(), self._overrides)])

    def set_parameters(self, replica=None, **params):
        """``model.parameters.<name> = value`` of the reference: re-evaluates all rate expressions."""
        targets = range(self.R) if replica is None else [replica]
        for r in targets:
            self._overrides[r].update(params)
        self.batch.set_rates(self._evaluate_rates())
        lut = self._evaluate_lut()
        if lut is not None:
            self.batch.set_otf_lut(lut)

    def get_parameters(self, replica=0):
        out = {k: v["value"] for k, v in self.ir["parameters"].items()}
        out.update(self._overrides[replica])
        return out

    @property
    def rate_constants(self):
        """{process: rate constant} of replica 0 (Model_Rate_Constants, kmos/run/__init__.py:1712)."""
        r = self.batch.rates[0]
        return dict(zip(self.process_names, r.tolist()))

    def get_backend(self):
        return self.ir["backend"]

    # ---- stepping ------------------------------------------------------------------------------------------
    def do_steps(self, n=10000):
        """proclist.do_kmc_steps(n) on every replica (kmos/run/__init__.py:416-432)."""
        self.batch.do_steps(int(n))

    def get_next_kmc_step(self):
        """KMC_Model.get_next_kmc_step (kmos/run/__init__.py:1365-1368): (proc, site) of replica 0, 1-based."""
        proc, site = self.batch.get_next_kmc_step()
        return int(proc[0]), int(site[0])

    def run_proc_nr(self, proc, site):
        """KMC_Model.run_proc_nr (kmos/run/__init__.py:1357-1363): execute process `proc` on site number `site`
        (both 1-based) on every replica of the batch."""
        self.batch.run_proc_nr(proc, site)

    # ---- observables ---------------------------------------------------------------------------------------
    def _adjustable(self):
        return [k for k in sorted(self.ir["parameters"]) if self.ir["parameters"][k].get("adjustable")]

    def get_param_header(self):
        return " ".join(self._adjustable())

    def get_tof_header(self):
        return " ".join(self.tofs)

    def get_occupation_header(self):
        # species sorted by name = species id order; sites in settings.site_names order
        # settings.site_names are `<layer>_<site>`; species sorted by name
        return " ".join("%s_%s" % (sp, site) for sp in sorted(self.species_names) for site in self.site_names)

    def get_std_header(self):
        return "#%s %s %s kmc_time simulated_time kmc_steps\n" % (
            self.get_param_header(), self.get_tof_header(), self.get_occupation_header())

    def get_atoms_all(self, reset_time_overrun=True):
        """Batched ``get_atoms(geometry=False)``: arrays over replicas of everything the reference's Expando
        carries (kmos/run/__init__.py:772-850)."""
        b = self.batch
        a = _Expando()
        a.kmc_time = b.kmc_time
        a.kmc_step = b.kmc_step
        a.procstat = b.procstat
        a.integ_rates = b.integ_rates
        a.occupation = b.occupation
        a.params = [[float(self.get_parameters(r)[k]) for k in self._adjustable()] for r in range(self.R)]
        delta_t = a.kmc_time - self._time
        delta_steps = a.kmc_step - self._steps
        cells = float(np.prod(self.size))
        a.tof_data = self._tof_data.copy()
        a.tof_integ = self._tof_integ.copy()
        ok = (delta_steps != 0) & (delta_t != 0.0)
        if ok.any():
            dps = (a.procstat - self._procstat)[ok] / delta_t[ok, None] / cells
            dig = (a.integ_rates - self._integ)[ok] / delta_t[ok, None] / cells
            a.tof_data[ok] = dps @ self.tof_matrix.T
            a.tof_integ[ok] = dig @ self.tof_matrix.T
        overrun = (delta_steps != 0) & (delta_t == 0.0) & (a.kmc_time > 0)
        if overrun.any() and reset_time_overrun:
            t = a.kmc_time.copy()
            t[overrun] = 0.0
            b.set_kmc_time(t)
            a.tof_data[overrun] = 0.0
            a.tof_integ[overrun] = 0.0
        a.delta_t = delta_t
        self._procstat, self._integ = a.procstat.copy(), a.integ_rates.copy()
        self._time, self._steps = a.kmc_time.copy(), a.kmc_step.copy()
        self._tof_data, self._tof_integ = a.tof_data.copy(), a.tof_integ.copy()
        return a

    def get_atoms(self, geometry=False, reset_time_overrun=True, replica=0):
        if geometry:
            raise NotImplementedError("ASE geometry output is outside the accelerated path (viewer)")
        allr = self.get_atoms_all(reset_time_overrun)
        a = _Expando()
        for k, v in allr.__dict__.items():
            setattr(a, k, v[replica])
        a.calc = None
        return a

    def get_std_sampled_data_all(self, samples, sample_size, tof_method="integ"):
        """rows[R][n_fields] in the order of get_std_header (kmos/run/__init__.py:865-966), all replicas."""
        occs, tofs, delta_ts, step_ts = [], [], [], []
        b = self.batch
        t0, step0 = b.kmc_time, b.kmc_step
        self.get_atoms_all(reset_time_overrun=False)
        for _ in range(samples):
            self.do_steps(sample_size // samples)
            atoms = self.get_atoms_all(reset_time_overrun=False)
            delta_ts.append(atoms.delta_t)
            step_ts.append(b.kmc_time_step)
            occs.append(atoms.occupation.reshape(self.R, -1))
            if tof_method == "procrates":
                tofs.append(atoms.tof_data)
            elif tof_method == "integ":
                tofs.append(atoms.tof_integ)
            else:
                raise NotImplementedError('tof_method="%s" not supported. Can be either procrates or integ.'
                                          % tof_method)
        occs, tofs = np.asarray(occs), np.asarray(tofs)
        delta_ts, step_ts = np.asarray(delta_ts), np.asarray(step_ts)
        occs_mean = (occs * step_ts[:, :, None]).sum(0) / step_ts.sum(0)[:, None]
        tof_mean = (tofs * delta_ts[:, :, None]).sum(0) / delta_ts.sum(0)[:, None]
        t1, step1 = b.kmc_time, b.kmc_step
        params = np.asarray(atoms.params, dtype=float).reshape(self.R, -1)
        return np.hstack([params, tof_mean, occs_mean, (t1 - t0)[:, None], t1[:, None],
                          (step1 - step0)[:, None].astype(float)])

    def get_std_sampled_data(self, samples, sample_size, tof_method="integ", output="str", replica=0):
        row = tuple(self.get_std_sampled_data_all(samples, sample_size, tof_method)[replica])
        if output == "str":
            return (" ".join(["%.5e"] * len(row)) + "\n") % row
        if output == "dict":
            return dict(zip(self.get_std_header()[1:].split(), row))
        raise UserWarning("Output format %s not defined. I only know 'str' and 'dict'" % output)

    # ---- configuration ---------------------------------------------------------------------------------------
    def _get_configuration(self, replica=0):
        """[X, Y, Z, N] int8 array of species ids (kmos/run/__init__.py:1396-1409)."""
        X, Y, Z = (int(x) for x in self.size)
        N = self.model.spuck
        lat = self.batch.lattice[replica]
        return lat.reshape(Z, Y, X, N).transpose(2, 1, 0, 3).astype(np.int8)

    def _set_configuration(self, config, replica=0):
        X, Y, Z = (int(x) for x in self.size)
        N = self.model.spuck
        config = np.asarray(config)
        if config.shape != (X, Y, Z, N):
            print("Config shape %s does not match" % (config.shape,))
            print("with model shape %s." % [X, Y, Z, N])
            return
        flat = config.transpose(2, 1, 0, 3).reshape(-1).astype(np.int32)
        self.batch.set_configuration(flat, replica=replica)

    def put(self, site, new_species, replica=0):
        """Put `new_species` (name or id) on site [x, y, z, n] and re-adjust the book-keeping
        (kmos/run/__init__.py:1243-1283 + _adjust_database)."""
        if len(site) != 4:
            raise ValueError("put: site must be [x, y, z, n] (n = 1-based site type), as in the reference")
        x, y, z, n = (int(v) for v in site)
        if not 1 <= n <= self.model.spuck:
            raise ValueError("put: site type n=%d outside 1..%d" % (n, self.model.spuck))
        sp = self.species_names.index(new_species) if isinstance(new_species, str) else int(new_species)
        cfg = self._get_configuration(replica)
        cfg[x % self.size[0], y % self.size[1], z % self.size[2], n - 1] = sp
        self._set_configuration(cfg, replica)

    def dump_config(self, filename, replica=0):
        np.save("%s.npy" % filename if not filename.endswith(".npy") else filename, self._get_configuration(replica))

    def load_config(self, filename, replica=0):
        f = filename if filename.endswith(".npy") else "%s.npy" % filename
        self._set_configuration(np.load(f), replica)

    # ---- f2py-shaped shim ------------------------------------------------------------------------------------
    def _shim(self):
        b = self.batch
        nr = {name.lower(): i + 1 for i, name in enumerate(self.process_names)}
        sp = {name.lower(): i for i, name in enumerate(self.species_names)}
        self.base = _Namespace(
            get_kmc_time=lambda: float(b.kmc_time[0]), get_kmc_step=lambda: int(b.kmc_step[0]),
            get_kmc_time_step=lambda: float(b.kmc_time_step[0]),
            get_procstat=lambda i: int(b.procstat[0, i - 1]), get_integ_rate=lambda i: float(b.integ_rates[0, i - 1]),
            get_nrofsites=lambda i: int(b.nr_of_sites[0, i - 1]), get_rate=lambda i: float(b.rates[0, i - 1]),
            get_accum_rate=lambda i: float(b.accum_rates[0, i - 1]),
            set_rate_const=lambda i, r: b.set_rate_const(i, r, replica=0),
            get_avail_site=lambda proc, field, switch: int(b.avail_sites(0)[proc - 1, field - 1, switch - 1]),
            set_kmc_time=lambda t: b.set_kmc_time(np.where(np.arange(b.R) == 0, float(t), b.kmc_time)),
            save_system=lambda: b.save_system("%s.reload" % self.ir.get("fixture", "kmc_model"), 0),
            reload_system=lambda: b.reload_system("%s.reload" % self.ir.get("fixture", "kmc_model"), 0),
            update_accum_rate=lambda: None, is_allocated=lambda: True, null_species=-1,
            get_null_species=lambda: -1, get_volume=lambda: b.volume)
        self.lattice = _Namespace(
            system_size=self.size.copy(), spuck=self.model.spuck, model_dimension=self.ir["model_dimension"],
            default_layer=self.ir["default_layer"],
            get_species=lambda site: int(self._get_configuration(0)[site[0] % self.size[0], site[1] % self.size[1],
                                                                     site[2] % self.size[2], site[3] - 1]),
            calculate_lattice2nr=lambda site: int(self.model.spuck * ((site[0] % self.size[0]) + self.size[0] * (
                (site[1] % self.size[1]) + self.size[1] * (site[2] % self.size[2]))) + site[3]),
            calculate_nr2lattice=lambda nr: [((nr - 1) // self.model.spuck) % self.size[0],
                                             ((nr - 1) // self.model.spuck) // self.size[0] % self.size[1],
                                             ((nr - 1) // self.model.spuck) // (self.size[0] * self.size[1]),
                                             (nr - 1) % self.model.spuck + 1],
            deallocate_system=self.deallocate)
        self.proclist = _Namespace(
            do_kmc_steps=self.do_steps, do_kmc_step=lambda: self.do_steps(1), nr_of_proc=self.model.n_proc,
            get_next_kmc_step=self.get_next_kmc_step, run_proc_nr=self.run_proc_nr,
            nr_of_species=self.model.n_species, backend=self.ir["backend"],
            get_occupation=lambda: b.occupation[0], **dict(list(nr.items()) + list(sp.items())))
