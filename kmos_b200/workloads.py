"""Synthetic replica batches for the BASELINE configs: deterministic rate matrices rates[R][P].

Rates are *inputs* of the hot path.  For the RuO2 headline workload they come from the model's own rate
expressions (examples/render_co_oxidation_ruo2.py:102-367) evaluated on a T x p_CO grid by
kmos_b200.rates (ideal-gas stand-in for the JANAF chemical potentials, documented there); every grid
point is repeated over `seeds` replicas that differ only in their Philox seed.
"""
import numpy as np

from . import rates as rates_mod


def ruo2_grid(ir, n_T=16, n_p=16, seeds=64, T_range=(450.0, 650.0), p_range=(1e-2, 1e2), p_O2=1.0):
    """rates[n_T*n_p*seeds][36], group_of[R] (grid point index), grid description."""
    Ts = np.linspace(T_range[0], T_range[1], n_T)
    ps = np.logspace(np.log10(p_range[0]), np.log10(p_range[1]), n_p)
    # the whole T x p_CO grid in one vectorised pass per process (rates.model_rates_grid)
    TT, PP = np.meshgrid(Ts, ps, indexing="ij")
    points = rates_mod.model_rates_grid(ir, {"T": TT.reshape(-1), "p_COgas": PP.reshape(-1)},
                                        overrides={"p_O2gas": float(p_O2)}, mu=rates_mod.standin_mu)
    rates = np.repeat(points, seeds, axis=0)
    group_of = np.repeat(np.arange(len(points), dtype=np.int32), seeds)
    return rates, group_of, {"T": Ts.tolist(), "p_COgas": ps.tolist(), "p_O2gas": p_O2, "seeds": seeds}


def zgb_grid(ir, n_y=64, seeds=64):
    ys = 0.30 + 0.25 * (np.arange(n_y) / max(n_y - 1, 1))
    points = np.asarray([rates_mod.model_rates(ir, {"yCO": float(y)}, mu=rates_mod.standin_mu) for y in ys])
    return np.repeat(points, seeds, axis=0), np.repeat(np.arange(n_y, dtype=np.int32), seeds), {"yCO": ys.tolist()}


def rates_for(name, ir, R):
    """A rate matrix of exactly R rows for fixture `name`."""
    if name.startswith("ruo2"):
        seeds = max(R // 256, 1)
        r, _g, _d = ruo2_grid(ir, seeds=seeds)
    elif name.startswith("zgb"):
        seeds = max(R // 64, 1)
        r, _g, _d = zgb_grid(ir, seeds=seeds)
    elif ir.get("process_defs"):
        r = np.asarray([rates_mod.model_rates(ir, mu=rates_mod.standin_mu)])
    else:  # fixtures parsed from committed Fortran carry no rate expressions: unit rate constants
        r = np.ones((1, len(ir["procs"])))
    reps = -(-R // len(r))
    return np.ascontiguousarray(np.tile(r, (reps, 1))[:R])
