"""Host-side rate-constant preparation (inputs of the hot path, not part of it).

The reference evaluates rate expressions with ``kmos.evaluate_rate_expression``
(kmos/__init__.py:67-189): tokenise, substitute unit constants (kmos/units.py), ``m_<formula>``
masses (ASE), ``mu_<gas>`` chemical potentials (JANAF tables, downloaded from NIST by
kmos/species.py:47-96) and the model parameters, then ``eval``.  The GPU box has neither kmos, ASE
nor network, so this module restates just enough of that to turn the rate expressions stored in the
model fixtures into the ``rates[R][P]`` matrix that both the oracle and the CUDA engine consume.
Because both sides are fed the same numbers, parity does not depend on this file.

Chemical potentials: ``mu`` may be a callable ``(gas, T, p) -> eV``, a ``kmos.species``-compatible provider
(attribute ``<gas>`` with a ``mu(T, p)`` method), or None.  None means "what the reference would use": the
reference's own ``kmos.species`` when it is importable (JANAF tables), otherwise the closed-form stand-in
``standin_mu`` **with a warning** -- the synthetic benchmark workloads pass ``standin_mu`` explicitly (SURVEY
8d).  A gas without a table gives 0 with a warning, as in the reference (kmos/__init__.py:157-160).
"""
import math
import re
import tokenize
import warnings
from io import StringIO

# CODATA 2010 values as in kmos/units.py:26-36
UNITS = {
    "pi": 3.14159265358979323846,
    "c": 2.99792458e8,
    "h": 6.62606957e-34,
    "hbar": 1.054571726e-34,
    "eV": 1.602176565e-19,
    "kboltzmann": 1.3806488e-23,
    "umass": 1.660538921e-27,
    "angstrom": 1.0e-10,
    "bar": 1.0e5,
}
UNIT_KEYS = ["pi", "c", "h", "hbar", "eV", "kboltzmann", "umass", "angstrom", "bar"]

# ASE >= 3.13 (IUPAC 2016) atomic masses for the elements the example models use
ATOMIC_MASSES = {"H": 1.008, "C": 12.011, "N": 14.007, "O": 15.999}
ATOMIC_MASSES_LEGACY = {"H": 1.00794, "C": 12.0107, "N": 14.0067, "O": 15.9994}

RATE_ALIASES = {"beta": "(1/(kboltzmann*T))"}

_KB_EV = 8.6173324e-5

# mu0(T) ~ a + b*T*(1 - ln(T/298.15)) eV at 1 bar: crude ideal-gas-like shape, NOT JANAF.
_MU_STANDIN = {
    "COgas": (-0.0, -2.05e-3),
    "O2gas": (-0.0, -2.13e-3),
    "CO2gas": (-0.0, -2.22e-3),
    "H2gas": (-0.0, -1.35e-3),
}


class MuStandinWarning(UserWarning):
    pass


def standin_mu(name, T, p):
    """Chemical potential stand-in in eV: linear-in-T entropy term + k_B T ln p (T in K, p in bar).
    NOT the JANAF value the reference uses; a gas without an entry gives 0 with a warning, as the reference
    does for a species without a table.  Accepts numpy arrays for T and p (model_rates_grid)."""
    key = name if name in _MU_STANDIN else name + "gas"
    if key not in _MU_STANDIN:
        warnings.warn("No chemical-potential table for %s: setting it to zero" % name, MuStandinWarning)
        return 0.0
    a, b = _MU_STANDIN[key]
    if hasattr(T, "shape") or hasattr(p, "shape"):
        import numpy as np
        T, p = np.asarray(T, dtype=float), np.asarray(p, dtype=float)
        return a + b * (T - 298.15) * 0.5 + b * T * 0.5 * np.log(np.maximum(T, 1.0) / 298.15) \
            + _KB_EV * T * np.log(p)
    T = float(T)
    return a + b * (T - 298.15) * 0.5 + b * T * 0.5 * math.log(max(T, 1.0) / 298.15) \
        + _KB_EV * T * math.log(float(p))


def _string2symbols(s):
    out = []
    for sym, count in re.findall(r"([A-Z][a-z]?)(\d*)", s):
        out.extend([sym] * (int(count) if count else 1))
    return out


def _reference_species():
    """kmos.species of an importable reference installation (JANAF tables), or None."""
    try:
        from kmos import species  # noqa: WPS433 -- optional: only where the reference is installed
        return species
    except Exception:
        return None


def resolve_mu(mu):
    """-> callable (gas, T, p) -> eV for any accepted `mu` argument (see the module docstring)."""
    if callable(mu):
        return mu
    provider = mu if mu is not None else _reference_species()
    if provider is not None:
        def from_provider(name, T, p):
            if not hasattr(provider, name):
                warnings.warn("No JANAF table assigned for %s: setting chemical potential to zero" % name,
                              MuStandinWarning)
                return 0.0
            return float(getattr(provider, name).mu(T, p))
        return from_provider

    def warned(name, T, p):
        warnings.warn("mu_%s evaluated with the closed-form stand-in (kmos_b200.rates.standin_mu), not with the "
                      "JANAF tables of kmos.species: pass mu=<provider> for the reference's rate constants" % name,
                      MuStandinWarning, stacklevel=3)
        return standin_mu(name, T, p)
    return warned


def evaluate_rate_expression(rate_expr, parameters=None, mu=None, masses=ATOMIC_MASSES):
    """Mirror of kmos.evaluate_rate_expression (kmos/__init__.py:67-189).

    ``parameters``: {name: {"value": v}} or {name: v}.  Same textual substitution + eval so that the
    floating-point result is the one the reference computes given the same masses and mu.
    """
    parameters = parameters or {}
    pdict = {}
    for k, v in parameters.items():
        pdict[k] = v["value"] if isinstance(v, dict) else v
    if not rate_expr:
        return 0.0
    for old, new in RATE_ALIASES.items():
        rate_expr = rate_expr.replace(old, new)
    tokens = list(tokenize.generate_tokens(StringIO(rate_expr).readline))
    replaced = []
    for i, token, _, _, _ in tokens:
        if token in ["sqrt", "exp", "sin", "cos", "pi", "pow", "log"]:
            replaced.append((i, "math." + token))
        elif token in UNITS:
            replaced.append((i, str(UNITS[token])))
        elif token.startswith("m_"):
            species_name = "_".join(token.split("_")[1:])
            replaced.append((i, "%s" % sum(masses[s] for s in _string2symbols(species_name))))
        elif token.startswith("mu_"):
            species_name = "_".join(token.split("_")[1:])
            if "T" not in pdict:
                raise KeyError('Need "T" in parameters to evaluate chemical potential.')
            if "p_%s" % species_name not in pdict:
                raise KeyError('Need "p_%s" in parameters to evaluate chemical potential.' % species_name)
            replaced.append((i, repr(resolve_mu(mu)(species_name, pdict["T"], pdict["p_%s" % species_name]))))
        elif token in pdict:
            s = str(pdict[token])
            for unit in UNIT_KEYS:
                s = s.replace(unit, "%s" % UNITS[unit])
            replaced.append((i, s))
        else:
            replaced.append((i, token))
    expr = tokenize.untokenize(replaced)
    return float(eval(expr, {"__builtins__": {}, "math": math}))


def evaluate_rate_expression_grid(rate_expr, parameters, grid, mu=None, masses=ATOMIC_MASSES):
    """evaluate_rate_expression over a whole sweep at once: `grid` maps parameter names to numpy arrays (one
    entry per sweep point, broadcastable); every other parameter keeps its scalar value.  One tokenisation and
    one eval per expression instead of one per sweep point.  -> array of the grid's shape (float64).
    Same substitutions as the scalar routine; transcendental functions are numpy's, which may differ from
    libm's in the last bit."""
    import numpy as np
    pdict = {k: (v["value"] if isinstance(v, dict) else v) for k, v in (parameters or {}).items()}
    shape = np.broadcast(*[np.asarray(v) for v in grid.values()]).shape if grid else ()
    if not rate_expr:
        return np.zeros(shape)
    for old, new in RATE_ALIASES.items():
        rate_expr = rate_expr.replace(old, new)
    env = {"__builtins__": {}, "np": np}
    for k, v in grid.items():
        env["_g_" + k] = np.asarray(v, dtype=float)

    def value_of(name):
        return env["_g_" + name] if name in grid else float(evaluate_rate_expression(str(pdict[name]), parameters))

    replaced = []
    for i, token, _, _, _ in tokenize.generate_tokens(StringIO(rate_expr).readline):
        if token in ["sqrt", "exp", "sin", "cos", "log"]:
            replaced.append((i, "np." + token))
        elif token == "pow":
            replaced.append((i, "np.power"))
        elif token == "pi":
            replaced.append((i, str(UNITS["pi"])))
        elif token in UNITS:
            replaced.append((i, str(UNITS[token])))
        elif token.startswith("m_"):
            species_name = "_".join(token.split("_")[1:])
            replaced.append((i, "%s" % sum(masses[s] for s in _string2symbols(species_name))))
        elif token.startswith("mu_"):
            species_name = "_".join(token.split("_")[1:])
            f = resolve_mu(mu)
            T, p = value_of("T"), value_of("p_%s" % species_name)
            try:
                val = f(species_name, T, p)
            except TypeError:  # a provider without array support
                val = np.vectorize(lambda t_, p_: f(species_name, float(t_), float(p_)))(T, p)
            env["_mu_" + species_name] = val
            replaced.append((i, "_mu_" + species_name))
        elif token in grid:
            replaced.append((i, "_g_" + token))
        elif token in pdict:
            s = str(pdict[token])
            for unit in UNIT_KEYS:
                s = s.replace(unit, "%s" % UNITS[unit])
            replaced.append((i, s))
        else:
            replaced.append((i, token))
    out = eval(tokenize.untokenize(replaced), env)
    return np.broadcast_to(np.asarray(out, dtype=float), shape).copy()


def model_rates_grid(ir, grid, overrides=None, **kw):
    """rates[points][P]: model_rates for every point of a sweep in one pass per process (`grid`: parameter
    name -> array over the sweep points)."""
    import numpy as np
    params = {k: dict(v) for k, v in ir["parameters"].items()}
    for k, v in (overrides or {}).items():
        params.setdefault(k, {})["value"] = v
    by_name = {p["name"].lower(): p for p in ir["process_defs"]}
    n = np.broadcast(*[np.asarray(v) for v in grid.values()]).shape
    cols = []
    for name in ir["procs"]:
        pd = by_name[name.lower()]
        cols.append(evaluate_rate_expression_grid(pd["rate_constant"], params, grid, **kw) if pd["enabled"]
                    else np.zeros(n))
    return np.stack([np.asarray(c, dtype=float).reshape(-1) for c in cols], axis=1)


def model_rates(ir, overrides=None, **kw):
    """rates[P] for a model fixture: every process' rate_constant with parameter overrides applied."""
    params = {k: dict(v) for k, v in ir["parameters"].items()}
    for k, v in (overrides or {}).items():
        params.setdefault(k, {})["value"] = v
    by_name = {p["name"].lower(): p for p in ir["process_defs"]}
    out = []
    for name in ir["procs"]:
        pd = by_name[name.lower()]
        out.append(evaluate_rate_expression(pd["rate_constant"], params, **kw) if pd["enabled"] else 0.0)
    return out
