"""kmc_model -- the module ``kmos export`` builds with f2py, re-implemented over libkmos_b200.so.

The unmodified reference front-end does ``from kmc_model import base, lattice, proclist`` (and optionally
``proclist_constants``, ``proclist_pars``) and ``import kmc_settings`` (kmos/run/__init__.py:86-129).  With this
directory's parent (``kmos_b200/dropin``) on ``sys.path`` those imports resolve to this package: every f2py
entry point ``kmos.run.KMC_Model`` uses is a plain function here that forwards to a one-replica batch on the GPU
(kmos_b200.engine.Batch -> the C-ABI of include/kmos_b200.h).  ``kmc_settings.py`` stays the file the reference
exporter writes.

Which model?  The f2py module is compiled per model; this one reads the model's rule tables -- the
``model_tables.json`` that kmos_b200.export writes next to the exported Fortran -- from ``$KMOS_B200_MODEL``,
else from ``model_tables.json`` in the current directory or next to ``kmc_settings.py`` on ``sys.path``.

There is no CPU path: ``proclist.init`` raises if the CUDA library or a GPU is missing.
"""
from . import _runtime
from . import base, lattice, proclist, proclist_constants  # noqa: F401

_runtime.load()
if _runtime.RT.ir["backend"] == "otf":
    from . import proclist_pars  # noqa: F401
