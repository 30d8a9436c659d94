"""Batched front-end over the C-ABI: R replicas of one kmos model on one GPU.

Mirrors, batched, what ``kmos.run.KMC_Model`` does with the f2py module (kmos/run/__init__.py:155-340,
416-432, 680-850): set rate constants, initialise the lattice, ``do_steps(n)``, read kmc_time / procstat /
integ_rates / occupation, derive TOFs.  All arrays carry a leading replica axis.
"""
import ctypes as C
import os

import numpy as np

from . import capi, tables


class Model(object):
    """The rule tables of one exported kmos model (what ``kmos export`` would compile)."""

    def __init__(self, ir=None, blob=None, info=None):
        if blob is None:
            blob, info = tables.build_blob(ir)
        self.ir, self.blob, self.info = ir, np.ascontiguousarray(blob, dtype=np.int32), info
        L = capi.lib()
        h = C.c_void_p()
        capi.check(L.kmos_b200_model_create(self.blob, self.blob.size, C.byref(h)))
        self.h = h
        self.n_proc = L.kmos_b200_model_nproc(h)
        self.n_species = L.kmos_b200_model_nspecies(h)
        self.spuck = L.kmos_b200_model_spuck(h)
        self.lut_size = L.kmos_b200_model_lut_size(h)
        self.backend = int(self.blob[2])
        self.default_layer = int(self.blob[9])

    @classmethod
    def from_json(cls, path):
        return cls(ir=tables.load_ir(path))

    def close(self):
        if getattr(self, "h", None):
            capi.lib().kmos_b200_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch(object):
    def __init__(self, model, n_replicas, size, device=0, seeds=None, replica_ids=None, rates=None, lut=None,
                 kernel=capi.KERNEL_AUTO, init=True, layer=None, proclist="auto", lpr=None):
        """proclist: the model's exporter-generated CUDA proclist (kmos_b200.codegen): a module path, "auto"
        (default: the cached build of this model, compiled now if there is none and nvcc is on the PATH; models
        or geometries the generator declines stay on the table interpreter), "build" (generate + compile, errors
        are raised) or None (table interpreter only).  lpr: lanes per replica for "auto"/"build" (default: the
        generator's choice)."""
        self.L = capi.lib()
        self.model = model
        self.R = int(n_replicas)
        size3 = np.ones(3, dtype=np.int32)
        size = np.atleast_1d(np.asarray(size, dtype=np.int32))
        size3[:len(size)] = size
        self.size = size3
        h = C.c_void_p()
        capi.check(self.L.kmos_b200_batch_create(model.h, self.R, size3, int(device), C.byref(h)))
        self.h = h
        self.P = model.n_proc
        self.volume = self.L.kmos_b200_batch_volume(h)
        self.ncells = self.volume // model.spuck
        self.layer = model.default_layer if layer is None else int(layer)
        self.proclist = None
        if kernel == capi.KERNEL_GENERATED and proclist in (None, "auto"):
            proclist = "build"
        if proclist and model.backend == capi.BACKEND_LOCAL_SMART and model.ir is not None:
            self.attach_proclist(proclist, lpr)
        elif proclist and proclist not in ("auto", "build"):
            self.attach_proclist(proclist, lpr)
        if kernel != capi.KERNEL_AUTO:
            self.select_kernel(kernel)
        self.set_seeds(np.arange(self.R, dtype=np.uint64) if seeds is None else seeds, replica_ids)
        if rates is not None:
            self.set_rates(rates)
        if lut is not None:
            self.set_otf_lut(lut)
        if init:
            self.init_state()

    # ---- setup -----------------------------------------------------------------------------------------
    def attach_proclist(self, proclist="auto", lpr=None):
        """Attach the model's exporter-generated CUDA proclist (kmos_b200_batch_attach_proclist).  Returns the
        module path, or None when "auto" has nothing to attach: the generator declines the model, there is no
        cached build and no compiler, or the module declines this lattice geometry."""
        import shutil
        from . import codegen, devtables
        if proclist not in ("auto", "build"):
            capi.check(self.L.kmos_b200_batch_attach_proclist(self.h, str(proclist).encode()))
            self.proclist = proclist
            return proclist
        can_build = proclist == "build" or bool(shutil.which(os.environ.get("NVCC", "nvcc")))

        def module(width):
            path = codegen.find_built(self.model.ir, self.model.blob, lpr=width)
            if path is None and can_build:
                path = codegen.build(self.model.ir, self.model.blob, lpr=width)
            return path

        env = os.environ.get("KMOS_B200_GEN_LPR")
        try:
            an = codegen._flatten(self.model.ir)
        except devtables.Unsupported:
            if proclist == "build":
                raise
            return None
        widths = [int(lpr)] if lpr else ([int(env)] if env else codegen.lane_group_widths(an["nproc"]))
        # How many replicas a lane-group width keeps resident depends on the lattice: attach every candidate
        # width once, score it (codegen.lane_group_score), keep the best.
        best = None
        for w in widths:
            path = module(w)
            if path is None:
                continue
            rc = self.L.kmos_b200_batch_attach_proclist(self.h, str(path).encode())
            if rc != 0:
                if proclist == "build" and len(widths) == 1:
                    capi.check(rc)
                continue  # e.g. a lattice smaller than twice the interaction range: the interpreter handles it
            if len(widths) == 1:
                self.proclist = path
                return path
            info = np.zeros(12, dtype=np.int64)
            self.L.kmos_b200_select_kernel(self.h, capi.KERNEL_GENERATED)
            capi.check(self.L.kmos_b200_kernel_info(self.h, info))
            resident = int(info[1] * info[3])
            per_sm = min(resident, max(1, -(-self.R // max(int(info[4]), 1))))
            rounds = [len(r) for r in codegen._schedule(an, w)]
            score = codegen.lane_group_score(an["nproc"], codegen.expected_max_rounds(rounds, 32 // w), 32 // w, per_sm)
            capi.check(self.L.kmos_b200_batch_detach_proclist(self.h))
            if best is None or score > best[0]:
                best = (score, w, path)
        if best is None:
            return None
        capi.check(self.L.kmos_b200_batch_attach_proclist(self.h, str(best[2]).encode()))
        self.proclist = best[2]
        return best[2]

    def select_kernel(self, kind):
        capi.check(self.L.kmos_b200_select_kernel(self.h, int(kind)))

    def kernel_info(self):
        info = np.zeros(12, dtype=np.int64)
        capi.check(self.L.kmos_b200_kernel_info(self.h, info))
        keys = ("kernel", "replicas_per_cta", "smem_bytes_per_cta", "ctas_per_sm", "sm_count",
                "state_bytes_per_replica", "table_bytes", "grid", "lists_in_l2", "registers", "split_lists",
                "image_bytes_per_replica")
        d = dict(zip(keys, (int(x) for x in info)))
        d["kernel_name"] = {capi.KERNEL_GENERIC: "generic", capi.KERNEL_SMEM: "smem",
                            capi.KERNEL_WARP_HBM: "warp_hbm", capi.KERNEL_GENERATED: "generated",
                            capi.KERNEL_OTF_FAST: "otf_fast"}[d["kernel"]]
        return d

    def set_seeds(self, seeds, replica_ids=None):
        s = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, dtype=np.uint64), (self.R,)))
        ids = None
        if replica_ids is not None:
            ids = np.ascontiguousarray(replica_ids, dtype=np.uint32)
            assert ids.size == self.R
        capi.check(self.L.kmos_b200_set_seeds(self.h, s, ids.ctypes.data_as(C.c_void_p) if ids is not None else None))

    def set_rates(self, rates):
        r = np.asarray(rates, dtype=np.float64)
        if r.ndim == 1:
            r = np.broadcast_to(r, (self.R, self.P))
        r = np.ascontiguousarray(r)
        assert r.shape == (self.R, self.P)
        capi.check(self.L.kmos_b200_set_rates(self.h, r))
        self._rates_host = r  # keep the pinned-or-not source alive until the async copy is consumed

    def set_rate_const(self, proc, rate, replica=-1):
        capi.check(self.L.kmos_b200_set_rate_const(self.h, int(replica), int(proc), float(rate)))

    def set_otf_lut(self, lut):
        t = np.asarray(lut, dtype=np.float64)
        if t.ndim == 1:
            t = np.broadcast_to(t, (self.R, t.size))
        t = np.ascontiguousarray(t)
        assert t.shape == (self.R, self.model.lut_size)
        capi.check(self.L.kmos_b200_set_otf_lut(self.h, t))
        self._lut_host = t

    def init_state(self, layer=None):
        capi.check(self.L.kmos_b200_init_state(self.h, self.layer if layer is None else int(layer)))

    def set_configuration(self, species, replica=-1, layer=None):
        s = np.ascontiguousarray(species, dtype=np.int32)
        assert s.size == (self.volume if replica >= 0 else self.R * self.volume)
        capi.check(self.L.kmos_b200_set_configuration(self.h, int(replica), s.reshape(-1),
                                                      self.layer if layer is None else int(layer)))

    # ---- stepping --------------------------------------------------------------------------------------
    def do_steps(self, n):
        capi.check(self.L.kmos_b200_do_kmc_steps(self.h, int(n)))

    def get_next_kmc_step(self):
        """proclist.get_next_kmc_step for every replica -> (proc[R], site[R]), 1-based, nothing executed."""
        proc, site = np.zeros(self.R, np.int32), np.zeros(self.R, np.int32)
        capi.check(self.L.kmos_b200_get_next_kmc_step(self.h, proc, site))
        return proc, site

    def run_proc_nr(self, proc, site):
        """proclist.run_proc_nr(proc[r], site[r]) on every replica (scalars are broadcast; proc 0 skips)."""
        proc = np.ascontiguousarray(np.broadcast_to(np.asarray(proc, np.int32), (self.R,)))
        site = np.ascontiguousarray(np.broadcast_to(np.asarray(site, np.int32), (self.R,)))
        capi.check(self.L.kmos_b200_run_proc_nr(self.h, proc, site))

    def set_kmc_time(self, t):
        """base.set_kmc_time for every replica: t[R] (a scalar is broadcast)."""
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float64), (self.R,)))
        capi.check(self.L.kmos_b200_set_kmc_time(self.h, t))

    # ---- restart files (base.save_system / base.reload_system) -----------------------------------------
    def save_system(self, path, replica=0):
        """Write replica `replica` as a reference-format .reload file (kmos_b200/checkpoint.py)."""
        from . import checkpoint
        r = int(replica)
        state = dict(kmc_time=self.kmc_time[r], kmc_step=self.kmc_step[r], procstat=self.procstat[r],
                     nr_of_sites=self.nr_of_sites[r], rates=self.rates[r], integ_rates=self.integ_rates[r],
                     lattice=self.lattice[r], avail_sites=self.avail_sites(r))
        checkpoint.write_reload(path, state)

    def reload_system(self, path, replica=0):
        """Restore replica `replica` from a .reload file; stepping continues bit-identically (same Philox key)."""
        from . import checkpoint
        st = checkpoint.read_reload(path)
        if st["nr_of_proc"] != self.P or st["volume"] != self.volume:
            raise ValueError("reload file is for nr_of_proc=%d volume=%d" % (st["nr_of_proc"], st["volume"]))
        integ = np.ascontiguousarray(st.get("integ_rates", np.zeros(self.P)), dtype=np.float64)
        capi.check(self.L.kmos_b200_reload_replica(
            self.h, int(replica), np.ascontiguousarray(st["lattice"], np.int32),
            np.ascontiguousarray(st["avail_sites"], np.int32).reshape(-1),
            np.ascontiguousarray(st["nr_of_sites"], np.int32), np.ascontiguousarray(st["procstat"], np.int64),
            integ, float(st["kmc_time"]), int(st["kmc_step"])))

    def set_stream(self, cuda_stream):
        """Run on a caller-owned CUDA stream (int handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        capi.check(self.L.kmos_b200_batch_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def synchronize(self):
        capi.check(self.L.kmos_b200_synchronize(self.h))

    def timer_start(self):
        capi.check(self.L.kmos_b200_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double(0)
        capi.check(self.L.kmos_b200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    # ---- observables -----------------------------------------------------------------------------------
    def _get(self, name, shape, dtype):
        out = np.zeros(shape, dtype=dtype)
        capi.check(getattr(self.L, "kmos_b200_get_" + name)(self.h, out))
        return out

    kmc_time = property(lambda self: self._get("kmc_time", self.R, np.float64))
    kmc_time_step = property(lambda self: self._get("kmc_time_step", self.R, np.float64))
    kmc_step = property(lambda self: self._get("kmc_step", self.R, np.int64))
    procstat = property(lambda self: self._get("procstat", (self.R, self.P), np.int64))
    integ_rates = property(lambda self: self._get("integ_rates", (self.R, self.P), np.float64))
    nr_of_sites = property(lambda self: self._get("nr_of_sites", (self.R, self.P), np.int32))
    accum_rates = property(lambda self: self._get("accum_rates", (self.R, self.P), np.float64))
    rates = property(lambda self: self._get("rates", (self.R, self.P), np.float64))
    lattice = property(lambda self: self._get("lattice", (self.R, self.volume), np.int32))
    occupation = property(lambda self: self._get("occupation", (self.R, self.model.n_species, self.model.spuck),
                                                 np.float64))
    status = property(lambda self: self._get("status", self.R, np.int32))
    error_info = property(lambda self: self._get("error_info", (self.R, 5), np.int32))

    def avail_sites(self, replica):
        out = np.zeros((self.P, self.volume, 2), dtype=np.int32)
        capi.check(self.L.kmos_b200_get_avail_sites(self.h, int(replica), out))
        return out

    def tally_words(self):
        return self.L.kmos_b200_tally_words(self.h)

    def reduce_tallies(self, group_of=None, n_groups=1, dev_ptr=None, want_host=True):
        """Per-group sums (see include/kmos_b200.h).  dev_ptr: device address to write into (for NCCL)."""
        words = self.tally_words()
        g = None
        if group_of is not None:
            g = np.ascontiguousarray(group_of, dtype=np.int32)
            assert g.size == self.R
        host = np.zeros((n_groups, words)) if want_host else None
        capi.check(self.L.kmos_b200_reduce_tallies(
            self.h, g.ctypes.data_as(C.c_void_p) if g is not None else None, int(n_groups),
            C.c_void_p(dev_ptr) if dev_ptr else None,
            host.ctypes.data_as(C.c_void_p) if host is not None else None))
        return host

    def split_tally(self, t):
        P, nocc = self.P, self.model.n_species * self.model.spuck
        t = np.asarray(t)
        return {"procstat": t[..., :P], "integ_rates": t[..., P:2 * P],
                "occupation": t[..., 2 * P:2 * P + nocc], "kmc_time": t[..., 2 * P + nocc],
                "kmc_steps": t[..., 2 * P + nocc + 1], "n_replicas": t[..., 2 * P + nocc + 2]}

    def close(self):
        if getattr(self, "h", None):
            self.L.kmos_b200_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _ShardView(Batch):
    """A fleet's shard seen through the Batch interface (borrowed handle: the fleet destroys it)."""

    def __init__(self, model, handle, n_replicas, size3, layer):
        self.L = capi.lib()
        self.model = model
        self.R = int(n_replicas)
        self.size = size3
        self.h = handle
        self.P = model.n_proc
        self.volume = self.L.kmos_b200_batch_volume(handle)
        self.ncells = self.volume // model.spuck
        self.layer = layer
        self.proclist = None

    def close(self):
        self.h = None


class Fleet(object):
    """R replicas of one model dealt to several GPUs of THIS process (kmos_b200_fleet_*, csrc/kb_fleet.h): the
    single-process counterpart of one Batch per torchrun rank.  gpu_ids: one entry per shard (default: every
    visible GPU); shard k holds replicas parallel.shard_bounds(R, k, len(gpu_ids)).  Same observables as Batch, in
    global replica order; trajectories do not depend on the number of shards (Philox counters carry the global
    replica number).  ``shards[k]`` is a Batch view for everything the fleet API does not forward."""

    def __init__(self, model, n_replicas, size, gpu_ids=None, seeds=None, rates=None, lut=None,
                 kernel=capi.KERNEL_AUTO, init=True, layer=None, proclist="auto", lpr=None):
        self.L = capi.lib()
        self.model = model
        self.R = int(n_replicas)
        size3 = np.ones(3, dtype=np.int32)
        size = np.atleast_1d(np.asarray(size, dtype=np.int32))
        size3[:len(size)] = size
        self.size = size3
        if gpu_ids is None:
            gpu_ids = list(range(max(self.L.kmos_b200_device_count(), 1)))
        self.gpu_ids = np.ascontiguousarray(gpu_ids, dtype=np.int32)
        s = None
        if seeds is not None:
            s = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, dtype=np.uint64), (self.R,)))
        h = C.c_void_p()
        capi.check(self.L.kmos_b200_fleet_create(model.h, self.R, size3, s.ctypes.data_as(C.c_void_p) if s is not None
                                                 else None, self.gpu_ids, int(self.gpu_ids.size), C.byref(h)))
        self.h = h
        self.P = model.n_proc
        self.layer = model.default_layer if layer is None else int(layer)
        self.shards, self.bounds = [], []
        for k in range(self.L.kmos_b200_fleet_n_shards(h)):
            lo, n = C.c_int32(0), C.c_int32(0)
            bh = self.L.kmos_b200_fleet_shard(h, k, C.byref(lo), C.byref(n))
            self.shards.append(_ShardView(model, C.c_void_p(bh), n.value, size3, self.layer))
            self.bounds.append((lo.value, lo.value + n.value))
        self.volume = self.shards[0].volume
        self.ncells = self.shards[0].ncells
        if kernel == capi.KERNEL_GENERATED and proclist in (None, "auto"):
            proclist = "build"
        if proclist and (proclist not in ("auto", "build") or
                         (model.backend == capi.BACKEND_LOCAL_SMART and model.ir is not None)):
            for sh in self.shards:  # per shard: the best lane-group width depends on the shard's replica count
                sh.attach_proclist(proclist, lpr)
        if kernel != capi.KERNEL_AUTO:
            self.select_kernel(kernel)
        if rates is not None:
            self.set_rates(rates)
        if lut is not None:
            self.set_otf_lut(lut)
        if init:
            self.init_state()

    def select_kernel(self, kind):
        capi.check(self.L.kmos_b200_fleet_select_kernel(self.h, int(kind)))

    def kernel_info(self):
        return [sh.kernel_info() for sh in self.shards]

    def set_rates(self, rates):
        r = np.asarray(rates, dtype=np.float64)
        if r.ndim == 1:
            r = np.broadcast_to(r, (self.R, self.P))
        r = np.ascontiguousarray(r)
        assert r.shape == (self.R, self.P)
        capi.check(self.L.kmos_b200_fleet_set_rates(self.h, r))
        self._rates_host = r

    def set_otf_lut(self, lut):
        t = np.asarray(lut, dtype=np.float64)
        if t.ndim == 1:
            t = np.broadcast_to(t, (self.R, t.size))
        t = np.ascontiguousarray(t)
        assert t.shape == (self.R, self.model.lut_size)
        capi.check(self.L.kmos_b200_fleet_set_otf_lut(self.h, t))
        self._lut_host = t

    def init_state(self, layer=None):
        capi.check(self.L.kmos_b200_fleet_init_state(self.h, self.layer if layer is None else int(layer)))

    def do_steps(self, n):
        """Enqueue n kMC steps on every GPU; returns at once, the getters synchronise."""
        capi.check(self.L.kmos_b200_fleet_do_kmc_steps(self.h, int(n)))

    def synchronize(self):
        capi.check(self.L.kmos_b200_fleet_synchronize(self.h))

    def _get(self, name, shape, dtype):
        out = np.zeros(shape, dtype=dtype)
        capi.check(getattr(self.L, "kmos_b200_fleet_get_" + name)(self.h, out))
        return out

    kmc_time = property(lambda self: self._get("kmc_time", self.R, np.float64))
    kmc_step = property(lambda self: self._get("kmc_step", self.R, np.int64))
    status = property(lambda self: self._get("status", self.R, np.int32))
    procstat = property(lambda self: self._get("procstat", (self.R, self.P), np.int64))
    integ_rates = property(lambda self: self._get("integ_rates", (self.R, self.P), np.float64))
    nr_of_sites = property(lambda self: self._get("nr_of_sites", (self.R, self.P), np.int32))
    lattice = property(lambda self: self._get("lattice", (self.R, self.volume), np.int32))
    occupation = property(lambda self: self._get("occupation", (self.R, self.model.n_species, self.model.spuck),
                                                 np.float64))

    # ---- the rest of the Batch interface, forwarded shard by shard (kmos_b200_fleet_shard + the batch API) ----
    def _cat(self, name):
        return np.concatenate([getattr(sh, name) for sh in self.shards], axis=0)

    kmc_time_step = property(lambda self: self._cat("kmc_time_step"))
    accum_rates = property(lambda self: self._cat("accum_rates"))
    rates = property(lambda self: self._cat("rates"))
    error_info = property(lambda self: self._cat("error_info"))

    def _owner(self, replica):
        """(shard, local index) of global replica `replica`."""
        replica = int(replica)
        for sh, (lo, hi) in zip(self.shards, self.bounds):
            if lo <= replica < hi:
                return sh, replica - lo
        raise IndexError(replica)

    def _rows(self, a, dtype):
        """A per-replica argument (scalars are broadcast) cut into the shards' blocks."""
        a = np.asarray(a, dtype)
        a = np.broadcast_to(a, (self.R,) + a.shape[1:]) if a.ndim <= 1 else a
        assert a.shape[0] == self.R
        return [np.ascontiguousarray(a[lo:hi]) for lo, hi in self.bounds]

    def set_rate_const(self, proc, rate, replica=-1):
        if replica < 0:
            for sh in self.shards:
                sh.set_rate_const(proc, rate)
        else:
            sh, r = self._owner(replica)
            sh.set_rate_const(proc, rate, replica=r)

    def set_configuration(self, species, replica=-1, layer=None):
        if replica >= 0:
            sh, r = self._owner(replica)
            return sh.set_configuration(species, replica=r, layer=layer)
        s = np.asarray(species, dtype=np.int32).reshape(self.R, self.volume)
        for sh, (lo, hi) in zip(self.shards, self.bounds):
            sh.set_configuration(s[lo:hi], layer=layer)

    def get_next_kmc_step(self):
        got = [sh.get_next_kmc_step() for sh in self.shards]
        return np.concatenate([g[0] for g in got]), np.concatenate([g[1] for g in got])

    def run_proc_nr(self, proc, site):
        for sh, p, q in zip(self.shards, self._rows(proc, np.int32), self._rows(site, np.int32)):
            sh.run_proc_nr(p, q)

    def set_kmc_time(self, t):
        for sh, x in zip(self.shards, self._rows(t, np.float64)):
            sh.set_kmc_time(x)

    def save_system(self, path, replica=0):
        sh, r = self._owner(replica)
        sh.save_system(path, r)

    def reload_system(self, path, replica=0):
        sh, r = self._owner(replica)
        sh.reload_system(path, r)

    def avail_sites(self, replica):
        sh, r = self._owner(replica)
        return sh.avail_sites(r)

    def tally_words(self):
        return self.shards[0].tally_words()

    def reduce_tallies(self, group_of=None, n_groups=1):
        """Per-group sums over all shards (layout: include/kmos_b200.h, split_tally)."""
        g = None
        if group_of is not None:
            g = np.ascontiguousarray(group_of, dtype=np.int32)
            assert g.size == self.R
        host = np.zeros((n_groups, self.tally_words()))
        capi.check(self.L.kmos_b200_fleet_reduce_tallies(
            self.h, g.ctypes.data_as(C.c_void_p) if g is not None else None, int(n_groups), host))
        return host

    def split_tally(self, t):
        return self.shards[0].split_tally(t)

    def close(self):
        if getattr(self, "h", None):
            for sh in self.shards:
                sh.close()
            self.L.kmos_b200_fleet_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measure_smem_bandwidth(device=0):
    """(GB/s, SM clock MHz) of the shared-memory streaming microbenchmark."""
    gb, mhz = C.c_double(0), C.c_double(0)
    capi.check(capi.lib().kmos_b200_measure_smem_bandwidth(int(device), C.byref(gb), C.byref(mhz)))
    return gb.value, mhz.value
