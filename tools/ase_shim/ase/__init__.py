"""Minimal stand-in for the parts of ASE that kmos touches at import/export time.

Only used by tools/make_fixtures.py in the authoring container (ASE is not installed there);
it never ships in the product path.  Masses are ASE's IUPAC-2016 values for the handful of
elements the kmos example models use.
"""
__version__ = "3.22.1"
from .atoms import Atoms  # noqa: F401
from . import atoms, data, io, symbols  # noqa: F401
