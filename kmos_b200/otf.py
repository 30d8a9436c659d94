"""otf backend: host-side tabulation of the on-the-fly rate expressions.

The reference's generated ``rate_<proc>(nr_vars)`` functions (kmos/io/__init__.py:2657-2964,
3000-3122) evaluate a user expression of ``rates(proc)``, ``userpar(i)``, ``chempots(i)`` and the
bystander counters ``nr_vars(k)`` every time a site is (re)registered.  The counters range over a small
finite domain (how many bystander sites carry the flag), so the whole function is tabulated once per
replica on the host and the device only looks values up (SURVEY 7 "FP reproducibility").  The same table
feeds the oracle and the CUDA engine, so parity does not depend on libm/pow details.

Integer powers follow gfortran: ``x**n`` with integer ``n`` is repeated multiplication (libgcc
``__powidf2``), everything else goes through libm.
"""
import ast
import itertools
import math
import re

import numpy as np

from .rates import evaluate_rate_expression, standin_mu


def _powi(x, m):
    n = -m if m < 0 else m
    y = x if n % 2 else 1.0
    n >>= 1
    while n:
        x = x * x
        if n % 2:
            y = y * x
        n >>= 1
    return 1.0 / y if m < 0 else y


def _fpow(a, b):
    if isinstance(b, (int, np.integer)) and not isinstance(b, bool):
        return _powi(a if isinstance(a, np.ndarray) else float(a), int(b))
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.power(a, b)
    return math.pow(a, b)


class _PowToCall(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Pow):
            return ast.copy_location(
                ast.Call(func=ast.Name(id="_fpow", ctx=ast.Load()), args=[node.left, node.right], keywords=[]), node)
        return node


def fortran_expr_to_callable(expr, proc_ids, userpar_ids, chempot_ids, constants, functions=None):
    """Translate a generated `rate_<proc>` right-hand side into f(rates, userpar, chempots, nr_vars).
    functions: replacements for the elementary functions (numpy's, for array arguments)."""
    py = expr
    py = re.sub(r"\brates\s*\(\s*(\w+)\s*\)", lambda m: "rates[%d]" % (proc_ids[m.group(1).lower()] - 1), py)
    py = re.sub(r"\buserpar\s*\(\s*(\w+)\s*\)", lambda m: "userpar[%d]" % userpar_ids[m.group(1)], py)
    py = re.sub(r"\bchempots\s*\(\s*(\w+)\s*\)", lambda m: "chempots[%d]" % chempot_ids[m.group(1)], py)
    py = re.sub(r"\bnr_vars\s*\(\s*(\d+)\s*\)", lambda m: "nr_vars[%d]" % (int(m.group(1)) - 1), py)
    py = re.sub(r"(\d)[dD]([-+]?\d)", r"\1e\2", py)
    tree = _PowToCall().visit(ast.parse(py.strip(), mode="eval"))
    ast.fix_missing_locations(tree)
    code = compile(tree, "<rate>", "eval")
    ns = {"__builtins__": {}, "_fpow": _fpow, "exp": math.exp, "sqrt": math.sqrt, "log": math.log,
          "sin": math.sin, "cos": math.cos, "abs": abs, "min": min, "max": max}
    for k, v in constants.items():
        ns[k] = float(v)
    if functions:
        ns.update(functions)

    def f(rates, userpar, chempots, nr_vars):
        env = dict(ns)
        env.update(rates=rates, userpar=userpar, chempots=chempots, nr_vars=nr_vars)
        out = eval(code, env)
        return out if functions else float(out)
    return f


def user_parameters(ir, overrides=None):
    """userpar(:) / chempots(:) values in the order of the generated proclist_pars module.  `overrides` maps
    parameter names to new values; an entry under a chemical potential's own name (``mu_<species>``, what
    proclist_pars.update_chempot receives from kmos.run.set_rate_constants) is taken as that potential."""
    params = {k: dict(v) for k, v in ir["parameters"].items()}
    given_mu = {}
    for k, v in (overrides or {}).items():
        if k in ir.get("chempots", []):
            given_mu[k] = float(v)
        else:
            params.setdefault(k, {})["value"] = v
    # first item of space-separated values: the reference does the same for its deprecated 'lattice_size'
    # parameter (kmos/run/__init__.py:2417-2428)
    userpar = [evaluate_rate_expression(str(params[name]["value"]).split(" ")[0], params)
               for name in ir.get("userpar", [])]
    chempots = []
    for name in ir.get("chempots", []):
        if name in given_mu:
            chempots.append(given_mu[name])
            continue
        species = name[len("mu_"):]
        chempots.append(standin_mu(species, evaluate_rate_expression(str(params["T"]["value"]), params),
                                   evaluate_rate_expression(str(params["p_" + species]["value"]), params)))
    return userpar, chempots


def build_lut_batch(ir, info, rates, overrides_list=None):
    """lut[R][lut_total] for R replicas at once: every table entry is evaluated once over all replicas (numpy
    arrays for rates[R][P] and the user parameters) instead of once per replica and entry."""
    rates = np.asarray(rates, dtype=float)
    R = rates.shape[0]
    lut = np.zeros((R, max(info["lut_total"], 1)))
    proc_ids = {p.lower(): i + 1 for i, p in enumerate(ir["procs"])}
    userpar_ids = {n: i for i, n in enumerate(ir.get("userpar", []))}
    chempot_ids = {n: i for i, n in enumerate(ir.get("chempots", []))}
    ups = [user_parameters(ir, (overrides_list[r] if overrides_list else None)) for r in
           (range(R) if overrides_list else range(1))]
    userpar = [np.array([u[0][k] for u in ups]) for k in range(len(ir.get("userpar", [])))]
    chempots = [np.array([u[1][k] for u in ups]) for k in range(len(ir.get("chempots", [])))]
    cols = [rates[:, q] for q in range(rates.shape[1])]
    by_lower = {k.lower(): v for k, v in ir["rate_expr"].items()}
    np_ns = {"exp": np.exp, "sqrt": np.sqrt, "log": np.log, "sin": np.sin, "cos": np.cos, "abs": np.abs,
             "min": np.minimum, "max": np.maximum}
    for name, g in info["gr"].items():
        expr = by_lower[name[len("gr_"):].lower()]
        f = fortran_expr_to_callable(expr, proc_ids, userpar_ids, chempot_ids, ir.get("pars_constants", {}),
                                     functions=np_ns)
        radix = g["radix"]
        for combo in itertools.product(*[range(r) for r in reversed(radix)]):
            nv = list(reversed(combo))
            idx, stride = 0, 1
            for k, r in enumerate(radix):
                idx += nv[k] * stride
                stride *= r
            lut[:, g["lut_offset"] + idx] = f(cols, userpar, chempots, nv)
    return lut


def rate_function(ir, proc, rates, overrides=None):
    """``proclist_pars.rate_<proc>`` as a Python callable of nr_vars (kmos/io/__init__.py:3000-3122): the user
    expression with the current rate constants, user parameters and chemical potentials bound."""
    proc_ids = {p.lower(): i + 1 for i, p in enumerate(ir["procs"])}
    userpar_ids = {n: i for i, n in enumerate(ir.get("userpar", []))}
    chempot_ids = {n: i for i, n in enumerate(ir.get("chempots", []))}
    userpar, chempots = user_parameters(ir, overrides)
    by_lower = {k.lower(): v for k, v in ir["rate_expr"].items()}
    f = fortran_expr_to_callable(by_lower[proc.lower()], proc_ids, userpar_ids, chempot_ids,
                                 ir.get("pars_constants", {}))
    rates = [float(x) for x in rates]
    return lambda nr_vars=(): f(rates, userpar, chempots, [int(v) for v in nr_vars])


def build_lut(ir, info, rates, overrides=None):
    """lut[lut_total] for one replica: every gr_<proc> over its whole nr_vars domain."""
    lut = np.zeros(max(info["lut_total"], 1))
    proc_ids = {p.lower(): i + 1 for i, p in enumerate(ir["procs"])}
    userpar_ids = {n: i for i, n in enumerate(ir.get("userpar", []))}
    chempot_ids = {n: i for i, n in enumerate(ir.get("chempots", []))}
    userpar, chempots = user_parameters(ir, overrides)
    by_lower = {k.lower(): v for k, v in ir["rate_expr"].items()}
    for name, g in info["gr"].items():
        expr = by_lower[name[len("gr_"):].lower()]
        f = fortran_expr_to_callable(expr, proc_ids, userpar_ids, chempot_ids, ir.get("pars_constants", {}))
        radix = g["radix"]
        for combo in itertools.product(*[range(r) for r in reversed(radix)]):
            nv = list(reversed(combo))
            idx, stride = 0, 1
            for k, r in enumerate(radix):
                idx += nv[k] * stride
                stride *= r
            lut[g["lut_offset"] + idx] = f(list(rates), userpar, chempots, nv)
    return lut
