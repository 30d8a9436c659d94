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


class ModelRunner(object):
    """Subclass and declare ModelParameter attributes, or pass ``parameters={name: ModelParameter}``."""

    def __init__(self, model, size=20, seeds=1, parameters=None, device=0, name=None):
        self.model_path, self.size, self.seeds, self.device = model, size, int(seeds), device
        self.runner_name = name or type(self).__name__
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

    def run(self, init_steps=1e5, sample_steps=1e5, samples=1, random_seed=1, outfile=None, per_replica=False):
        """Returns (header, rows).  One row per grid point (mean over its seeds) unless ``per_replica``."""
        points = self.grid_points()
        per_rep = [pt for pt in points for _ in range(self.seeds)]
        outfile = outfile or os.path.abspath("%s.dat" % self.runner_name)
        with KMC_Model(self.model_path, size=self.size, n_replicas=len(per_rep), parameters=per_rep,
                       device=self.device, random_seed=random_seed) as model:
            model.do_steps(int(init_steps))
            model.get_atoms_all()
            rows = model.get_std_sampled_data_all(samples, int(sample_steps), tof_method="integ")
            header = model.get_std_header()
            params_dump = "".join("# %s = %s\n" % (k, v) for k, v in sorted(model.get_parameters(0).items()))
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
