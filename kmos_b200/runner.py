"""Batched counterpart of ``kmos.run.ModelRunner`` (kmos/run/__init__.py:2005-2366).

The reference scans a regular parameter grid with a pool of processes, one ``KMC_Model`` per grid point:
``do_steps(init_steps)``, then ``get_std_sampled_data(samples, sample_steps, tof_method="integ")``, and appends
the row to ``<runner>.dat`` (header = ``get_std_header()``).  Here every grid point (times ``seeds`` replicas)
is one replica of a single GPU batch and the same rows are written in one go.

    class ScanKinetics(ModelRunner):
        p_O2gas = PressureParameter(1)
        T = TemperatureParameter(600)
        p_COgas = PressureParameter(min=1, max=10, steps=40)
    ScanKinetics("ruo2_local_smart.json", size=20).run(init_steps=1e5, sample_steps=1e5)
"""
import itertools
import os
from collections import OrderedDict

import numpy as np

from .model import KMC_Model


# ModelParameter and its four subclasses are an API mirror of kmos/run/__init__.py:1896-1985 (same constructor
# arguments, same __repr__, same grids) so that existing ModelRunner subclasses keep working unchanged.
class ModelParameter(object):
    """A variable to scan: ``min``, ``max``, ``steps`` (kmos/run/__init__.py:1896-1925)."""

    def __init__(self, min, max=None, steps=1, type=None, unit=""):
        self.min, self.max = min, (max if max is not None else min)
        self.steps, self.type, self.unit = steps, type, unit

    def __repr__(self):
        return "[%s] min: %s, max: %s, steps: %s" % (self.type, self.min, self.max, self.steps)

    def get_grid(self):
        return np.linspace(self.min, self.max, self.steps)


class PressureParameter(ModelParameter):
    """ln(p) regular (kmos/utils/__init__.py:1438-1446 p_grid)."""

    def __init__(self, *a, **k):
        k.update(type="pressure", unit="bar")
        super(PressureParameter, self).__init__(*a, **k)

    def get_grid(self):
        return np.logspace(np.log10(self.min), np.log10(self.max), self.steps)


class TemperatureParameter(ModelParameter):
    """1/T regular (kmos/utils/__init__.py:1420-1435 T_grid)."""

    def __init__(self, *a, **k):
        k.update(type="temperature", unit="K")
        super(TemperatureParameter, self).__init__(*a, **k)

    def get_grid(self):
        grid = list(np.linspace(self.max ** -1.0, self.min ** -1.0, self.steps))
        grid.reverse()
        return np.array([x ** -1.0 for x in grid])


class LogParameter(ModelParameter):
    def __init__(self, *a, **k):
        k.update(type="log")
        super(LogParameter, self).__init__(*a, **k)

    def get_grid(self):
        return np.logspace(self.min, self.max, self.steps)


class LinearParameter(ModelParameter):
    def __init__(self, *a, **k):
        k.update(type="linear")
        super(LinearParameter, self).__init__(*a, **k)


def _run_points(model_path, size, seeds, points, first_point, device, init_steps, sample_steps, samples,
                random_seed, model_factory=None, gpu_ids=None):
    """The rows of `points` (grid points first_point, first_point+1, ... of the whole scan), `seeds` replicas
    each, as one batch on one GPU.  Philox keys follow the *global* replica index, so a scan gives the same
    rows however its points are dealt to GPUs.  gpu_ids: the whole scan (first_point 0) as one fleet of this
    process (engine.Fleet) instead.  -> (header, rows[len(points)*seeds][fields], parameter dump)"""
    per_rep = [pt for pt in points for _ in range(seeds)]
    gid = first_point * seeds + np.arange(len(per_rep))
    keys = (np.uint64(random_seed) + gid.astype(np.uint64))
    factory = model_factory or KMC_Model
    if gpu_ids is not None:
        assert first_point == 0
        where = dict(gpu_ids=list(gpu_ids))
    else:
        where = dict(device=device, replica_ids=gid.astype(np.uint32))
    with factory(model_path, size=size, n_replicas=len(per_rep), parameters=per_rep, seeds=keys, **where) as model:
        model.do_steps(int(init_steps))
        model.get_atoms_all()
        rows = model.get_std_sampled_data_all(samples, int(sample_steps), tof_method="integ")
        header = model.get_std_header()
        params_dump = "".join("# %s = %s\n" % (k, v) for k, v in sorted(model.get_parameters(0).items()))
    return header, np.asarray(rows), params_dump


def _spawn_worker(conn, args):
    try:
        conn.send(("ok", _run_points(*args)))
    except Exception as e:  # the parent re-raises
        import traceback
        conn.send(("error", "%s\n%s" % (e, traceback.format_exc())))
    finally:
        conn.close()


class ModelRunner(object):
    """Subclass and declare ModelParameter attributes, or pass ``parameters={name: ModelParameter}``.

    API mirror of kmos.run.ModelRunner (kmos/run/__init__.py:2005-2366): the reference's ``run(cores=N)`` deals
    grid points to a pool of N processes that append to one ``.dat`` file under a lock file; ``run(gpus=N)``
    deals contiguous blocks of grid points (kmos_b200.parallel.shard_bounds) to N GPUs of the box -- one
    process per GPU, no traffic while stepping -- gathers the rows on rank 0 and writes the same file;
    ``run(gpu_ids=[...])`` keeps one process and drives the GPUs through a fleet (kmos_b200_fleet_*)."""

    def __init__(self, model, size=20, seeds=1, parameters=None, device=0, name=None, model_factory=None):
        self.model_path, self.size, self.seeds, self.device = model, size, int(seeds), device
        self.runner_name = name or type(self).__name__
        self.model_factory = model_factory
        self.parameters = OrderedDict()
        for klass in reversed(type(self).__mro__):
            for key, item in vars(klass).items():
                if isinstance(item, ModelParameter):
                    self.parameters[key] = item
        for key, item in (parameters or {}).items():
            self.parameters[key] = item

    def grid_points(self):
        grids = [p.get_grid() for p in self.parameters.values()]
        return [dict(zip(self.parameters.keys(), (float(v) for v in pt))) for pt in itertools.product(*grids)]

    def _gather(self, points, gpus, args, gpu_ids=None):
        """-> (header, rows of every grid point in grid order, parameter dump), or None on ranks > 0."""
        from . import parallel
        try:
            import torch.distributed as dist
            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        except ImportError:
            distributed = False
        if distributed:  # launched by torch.distributed.run: this process is one rank with its own GPU
            rank, world = dist.get_rank(), dist.get_world_size()
            lo, hi = parallel.shard_bounds(len(points), rank, world)
            device = int(os.environ.get("LOCAL_RANK", rank))
            mine = _run_points(self.model_path, self.size, self.seeds, points[lo:hi], lo, device, *args,
                               model_factory=self.model_factory) if hi > lo else None
            parts = [None] * world if rank == 0 else None
            dist.gather_object(mine, parts, dst=0)
            if rank != 0:
                return None
        elif gpu_ids is not None:  # this process drives the listed GPUs itself (kmos_b200_fleet_*)
            parts = [_run_points(self.model_path, self.size, self.seeds, points, 0, self.device, *args,
                                 model_factory=self.model_factory, gpu_ids=gpu_ids)]
        elif gpus and gpus > 1:  # one worker process per GPU of this box
            import multiprocessing as mp
            ctx = mp.get_context("spawn")
            jobs = []
            for rank in range(gpus):
                lo, hi = parallel.shard_bounds(len(points), rank, gpus)
                if hi == lo:
                    continue
                parent, child = ctx.Pipe()
                wargs = (self.model_path, self.size, self.seeds, points[lo:hi], lo, rank) + tuple(args) + \
                    (self.model_factory,)
                proc = ctx.Process(target=_spawn_worker, args=(child, wargs))
                proc.start()
                jobs.append((proc, parent))
            parts = []
            for proc, parent in jobs:
                status, payload = parent.recv()
                proc.join()
                if status != "ok":
                    raise RuntimeError("ModelRunner worker failed: %s" % payload)
                parts.append(payload)
        else:
            parts = [_run_points(self.model_path, self.size, self.seeds, points, 0, self.device, *args,
                                 model_factory=self.model_factory)]
        parts = [p for p in parts if p is not None]
        return parts[0][0], np.vstack([p[1] for p in parts]), parts[0][2]

    def run(self, init_steps=1e5, sample_steps=1e5, samples=1, random_seed=1, outfile=None, per_replica=False,
            gpus=None, gpu_ids=None):
        """Returns (header, rows).  One row per grid point (mean over its seeds) unless ``per_replica``.
        gpus: number of GPUs of this box to deal the grid points to (default: one, ``device``); under
        torch.distributed.run the ranks of the job are used instead and only rank 0 returns rows.
        gpu_ids: run the scan as one fleet on these GPUs from this process, no workers (same rows again)."""
        points = self.grid_points()
        outfile = outfile or os.path.abspath("%s.dat" % self.runner_name)
        got = self._gather(points, gpus, (init_steps, sample_steps, samples, random_seed), gpu_ids)
        if got is None:
            return None, None
        header, rows, params_dump = got
        if not per_replica:
            rows = rows.reshape(len(points), self.seeds, -1).mean(axis=1)
        new = not os.path.exists(outfile)
        with open(outfile, "a") as out:
            if new:
                out.write(header)
                out.write(params_dump)
                out.write("# If one or more parameters change between data lines\n"
                          "# the set above corresponds to the first line.\n")
            for row in rows:
                out.write((" ".join(["%.5e"] * len(row)) + "\n") % tuple(row))
        return header, rows
