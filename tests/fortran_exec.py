"""A second, parser-independent execution of the generated Fortran (test infrastructure).

The oracle and the CUDA engines all run byte-code that kmos_b200/fortran_ir.py + tables.py derive from the
text `kmos export` writes.  A mis-read nli_* decision tree or a wrongly ordered del/add list would pass every
oracle-vs-GPU test.  This module does not use those two files: it reads the same export directory
(run_proc_*.f90, nli_*.f90, proclist.f90, proclist_pars.f90, proclist_constants.f90, lattice.f90) and
*interprets the Fortran statements themselves* -- ``select case``, ``if (can_do(..)) then``, ``call add_proc /
del_proc / replace_species / update_rates_matrix``, function results, the ``nr_vars`` counters of ``gr_<proc>``
and the arithmetic of ``rate_<proc>`` -- on a plain-Python restatement of the base module, under the shared
Philox stream.  tests/test_fortran_exec.py compares it with the oracle event by event.

Restated from the reference (the model-independent part; citations are to /root/reference):
  add_proc / del_proc / can_do      kmos/fortran_src/base.mpy:211-321, base_lat_int.mpy:249,299 (proc 0 is a
                                    no-op), base_otf.f90:221-364 (rates_matrix book-keeping)
  update_accum_rate                 base.mpy:603-623, base_otf.f90:687-717
  determine_procsite                base.mpy:1075-1120, base_otf.f90:1213-1277
  interval_search_real              base.mpy:1234-1338
  update_clocks                     base.mpy:1123-1161
  lattice2nr / get_species          kmos/fortran_src/lattice.mpy:146-210
  do_kmc_steps                      kmos/fortran_src/proclist_generic_subroutines.mpy:1-44
"""
import ast
import glob
import math
import os
import re

from kmos_b200.otf import _PowToCall, _fpow  # gfortran's x**n for integer n; arithmetic only, no parsing


class Vec(tuple):
    """integer, dimension(4): `cell + (/1, 0, 0, 1/)`, `site(3)`."""

    def __add__(self, other):
        return Vec(a + b for a, b in zip(self, other))

    __radd__ = __add__

    def __call__(self, i):
        return self[i - 1]


class Arr(list):
    """1-based Fortran array: `nr_vars(2)`, `userpar(j_co_co)`, `rates(co_ads)`."""

    def __call__(self, i):
        return self[i - 1]


class _Return(Exception):
    pass


def _logical_lines(text):
    out, cur = [], ""
    for raw in text.splitlines():
        line = raw
        # strip comments (no '!' occurs inside the strings of the generated code we execute)
        if "!" in line:
            q = line.find("!")
            if line.count('"', 0, q) % 2 == 0 and line.count("'", 0, q) % 2 == 0:
                line = line[:q]
        line = line.strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1].rstrip() + " "
            continue
        cur += line
        out.append(cur)
        cur = ""
    if cur:
        out.append(cur)
    return out


def _split_statements(line):
    parts, depth, cur = [], 0, ""
    for ch in line:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == ";" and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


_DECL = re.compile(r"^(integer|real|character|logical|use\b|implicit|private|public|contains|module\b|end module|"
                   r"intent|allocate|deallocate|print|stop)")
_REL = [(r"\.ne\.", "!="), (r"\.eq\.", "=="), (r"\.ge\.", ">="), (r"\.le\.", "<="), (r"\.gt\.", ">"), (r"\.lt\.", "<"),
        (r"\.and\.", " and "), (r"\.or\.", " or "), (r"\.not\.", " not ")]


def _expr(src):
    """Fortran expression text -> compiled Python expression."""
    s = src.strip()
    s = s.replace("(/", " Vec((").replace("/)", ",)) ")
    for a, b in _REL:
        s = re.sub(a, b, s)
    s = re.sub(r"(\d)[dD]([-+]?\d)", r"\1e\2", s)   # 1.d0 -> 1.e0
    s = re.sub(r",\s*:\s*\)", ", None)", s)          # nr2lattice(n, :)
    tree = _PowToCall().visit(ast.parse(s.strip(), mode="eval"))
    ast.fix_missing_locations(tree)
    return compile(tree, "<f90>", "eval")


def _matching_paren(s, i):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses: %r" % s)


def _split_args(s):
    args, depth, cur = [], 0, ""
    i = 0
    while i < len(s):
        ch = s[i]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            args.append(cur.strip())
            cur = ""
        else:
            cur += ch
        i += 1
    if cur.strip():
        args.append(cur.strip())
    return args


class Routine(object):
    def __init__(self, name, kind, params, body):
        self.name, self.kind, self.params, self.body = name, kind, params, body
        self.local_arrays = {}


def _parse_block(stmts, pos, enders):
    """-> (list of statement nodes, index of the ender that stopped the block)."""
    out = []
    while pos < len(stmts):
        s = stmts[pos]
        low = s
        if any(re.match(e, low) for e in enders):
            return out, pos
        m = re.match(r"^select\s*case\s*\(", low)
        if m:
            j = _matching_paren(low, m.end() - 1)
            sel = _expr(low[m.end():j])
            pos += 1
            cases = []
            while not re.match(r"^end\s*select", stmts[pos]):
                c = stmts[pos]
                if re.match(r"^case\s+default", c):
                    keys = None
                else:
                    mc = re.match(r"^case\s*\(", c)
                    assert mc, "expected case: %r" % c
                    keys = [_expr(k) for k in _split_args(c[mc.end():_matching_paren(c, mc.end() - 1)])]
                body, pos = _parse_block(stmts, pos + 1, [r"^case\b", r"^end\s*select"])
                cases.append((keys, body))
            out.append(("select", sel, cases))
            pos += 1
            continue
        m = re.match(r"^if\s*\(", low)
        if m:
            j = _matching_paren(low, m.end() - 1)
            cond = _expr(low[m.end():j])
            rest = low[j + 1:].strip()
            if rest == "then":
                body, pos = _parse_block(stmts, pos + 1, [r"^else\b", r"^end\s*if"])
                orelse = []
                if re.match(r"^else\b", stmts[pos]):
                    assert stmts[pos].strip() == "else", "else if is not generated: %r" % stmts[pos]
                    orelse, pos = _parse_block(stmts, pos + 1, [r"^end\s*if"])
                out.append(("if", cond, body, orelse))
                pos += 1
            else:
                inner, _ = _parse_block([rest], 0, [])
                out.append(("if", cond, inner, []))
                pos += 1
            continue
        m = re.match(r"^do\s+(\w+)\s*=\s*(.+)$", low)
        if m:
            lo_hi = _split_args(m.group(2))
            body, pos = _parse_block(stmts, pos + 1, [r"^end\s*do"])
            out.append(("do", m.group(1), _expr(lo_hi[0]), _expr(lo_hi[1]), body))
            pos += 1
            continue
        m = re.match(r"^call\s+(\w+)\s*(\(.*\))?$", low)
        if m:
            args = _split_args(m.group(2)[1:-1]) if m.group(2) else []
            out.append(("call", m.group(1), [_expr(a) for a in args if not a.startswith("put=")]))
            pos += 1
            continue
        if low == "return":
            out.append(("return",))
            pos += 1
            continue
        if _DECL.match(low):
            out.append(("decl", low))
            pos += 1
            continue
        m = re.match(r"^(\w+)\s*(\(([^=]*)\))?\s*=(?!=)\s*(.+)$", low)
        if m:
            idx = m.group(3)
            out.append(("assign", m.group(1), None if idx is None else (":" if idx.strip() == ":" else _expr(idx)),
                        _expr(m.group(4))))
            pos += 1
            continue
        raise ValueError("statement not understood: %r" % s)
    return out, pos


class FortranModel(object):
    """The routines and constants of one export directory, as parsed statement trees."""

    def __init__(self, path):
        self.path = path
        self.constants = {}
        self.routines = {}
        files = sorted(glob.glob(os.path.join(path, "*.f90")))
        skip = ("base.f90", "main.f90", "kind_values.f90", "kind_values_f2py.f90", "f2py_selected_kind.f90", "base_acf.f90")
        for f in files:
            if os.path.basename(f) in skip:
                continue
            with open(f) as fh:
                lines = [ln.lower() for ln in _logical_lines(fh.read())]
            # lattice.f90: constants only -- its routines wrap the base module, which is restated natively below
            self._scan(lines, os.path.basename(f), constants_only=os.path.basename(f) == "lattice.f90")
        names = set(self.routines)
        self.backend = "otf" if any(n.startswith("gr_") for n in names) else \
            ("lat_int" if any(n.startswith("nli_") for n in names) else "local_smart")
        self.nr_of_proc = self.constants["nr_of_proc"]
        self.spuck = self.constants["spuck"]
        self.proc_names = self._process_names()

    def _scan(self, lines, fname, constants_only=False):
        i = 0
        stmts = []
        for ln in lines:
            stmts += _split_statements(ln)
        while i < len(stmts):
            s = stmts[i]
            m = re.match(r"^(?:pure\s+|recursive\s+)*(subroutine|function)\s+(\w+)\s*(\(([^)]*)\))?", s)
            if m:
                kind, name = m.group(1), m.group(2)
                params = [p.strip() for p in (m.group(4) or "").split(",") if p.strip()]
                j = i + 1
                while not re.match(r"^end\s*(subroutine|function)", stmts[j]):
                    j += 1
                body_src = stmts[i + 1:j]
                r = Routine(name, kind, params, None)
                for d in body_src:
                    md = re.match(r"^(integer|real)[^:]*dimension\s*\(\s*(\d+)\s*\)[^:]*::\s*(.+)$", d)
                    if md:
                        for v in md.group(3).split(","):
                            r.local_arrays[v.strip()] = int(md.group(2))
                r.src = body_src
                if not constants_only:
                    self.routines[name] = r
                i = j + 1
                continue
            m = re.match(r"^(integer|real)[^:]*::\s*(\w+)\s*=\s*([-+]?[\w.]+(?:[-+]\d+)?)$", s)
            if m and "dimension" not in s:
                try:
                    self.constants[m.group(2)] = int(m.group(3))
                except ValueError:
                    try:
                        self.constants[m.group(2)] = float(re.sub(r"[dD]", "e", m.group(3)))
                    except ValueError:
                        if m.group(3) in self.constants:
                            self.constants[m.group(2)] = self.constants[m.group(3)]
            i += 1

    def routine(self, name):
        r = self.routines[name]
        if r.body is None:
            r.body, _ = _parse_block(r.src, 0, [])
        return r

    def _process_names(self):
        """process number -> name, from run_proc_nr's `case(<name>)` labels (proclist.f90)."""
        out = {}
        for s in self.routines["run_proc_nr"].src:
            m = re.match(r"^case\s*\(([\w\s,]+)\)$", s)   # lat_int: one case lists a whole process group
            for label in (m.group(1).split(",") if m else []):
                label = label.strip()
                if label in self.constants:
                    out[self.constants[label]] = label
        return out

    def userpar_names(self):
        """names of userpar(:) in index order (proclist_pars.f90: `integer, public :: <name> = <index>`)."""
        return [n for n, _i in self.index_declarations()[0]]

    def index_declarations(self):
        """([(name, index)] of userpar(:), [(name, index)] of chempots(:)): the index constants of
        proclist_pars.f90 in declaration order; the chemical potentials restart at 1."""
        with open(os.path.join(self.path, "proclist_pars.f90")) as fh:
            text = [ln.lower() for ln in _logical_lines(fh.read())]
        groups = [[]]
        for ln in text:
            if ln.startswith("contains"):
                break
            m = re.match(r"^integer\(kind=iint\), public :: (\w+) = (\d+)$", ln)
            if m:
                if groups[-1] and int(m.group(2)) <= groups[-1][-1][1]:
                    groups.append([])
                groups[-1].append((m.group(1), int(m.group(2))))
        return groups[0], (groups[1] if len(groups) > 1 else [])


class Executor(object):
    """Base-module state + interpreter for one replica."""

    def __init__(self, model, size, rates, seed, replica, philox_step, userpar=None, chempots=None, layer=None):
        self.m = model
        self.size = list(size) + [1] * (3 - len(size))
        self.spuck = model.spuck
        self.volume = self.size[0] * self.size[1] * self.size[2] * self.spuck
        P = model.nr_of_proc
        self.P = P
        self.otf = model.backend == "otf"
        self.lattice = [0] * self.volume
        self.avail1 = [[0] * (self.volume + 1) for _ in range(P + 1)]   # avail_sites(proc, k, 1)
        self.avail2 = [[0] * (self.volume + 1) for _ in range(P + 1)]   # avail_sites(proc, site, 2)
        self.nr_of_sites = [0] * (P + 1)
        self.rates = Arr(float(x) for x in rates)
        self.accum_rates = [0.0] * (P + 1)
        self.procstat = [0] * (P + 1)
        self.rates_matrix = [[0.0] * (self.volume + 2) for _ in range(P + 1)] if self.otf else None
        self.kmc_time = 0.0
        self.kmc_step = 0
        self.seed, self.replica, self.philox_step = seed, replica, philox_step
        # a model that declares a null species hands it to base.set_null_species (kmos/run/__init__.py:232-240)
        self.null_species = model.constants.get("null_species", -1)
        self.g = dict(model.constants)
        self.g.update(Vec=Vec, _fpow=_fpow, exp=math.exp, sqrt=math.sqrt, log=math.log, abs=abs, min=min, max=max,
                      real=float, int=int,
                      get_species=self.get_species, can_do=self.can_do, lattice2nr=self.lattice2nr,
                      avail_sites=self.avail_sites_ref, nr2lattice=self.nr2lattice, rates=self.rates,
                      system_size=Arr(self.size), null_species=self.null_species,
                      userpar=Arr(userpar or []), chempots=Arr(chempots or []))
        for name, r in model.routines.items():
            if r.kind == "function":
                self.g[name] = self._make_function(name)
        self.builtin = {"add_proc": self.add_proc, "del_proc": self.del_proc, "replace_species": self.replace_species,
                        "update_rates_matrix": self.update_rates_matrix, "reset_site": self.reset_site,
                        "increment_procstat": self.increment_procstat, "random_seed": lambda *a: None}
        layer_id = model.constants[layer] if isinstance(layer, str) else layer
        if layer_id is None:
            layer_id = model.constants.get("default_layer", 0)
        self.call("initialize_state", [layer_id, 0])

    # ---- lattice module (lattice.mpy:146-210) ------------------------------------------------------------
    def lattice2nr(self, *a):
        x, y, z, n = a[0] if len(a) == 1 else a
        Lx, Ly, Lz = self.size
        return ((x % Lx) + Lx * ((y % Ly) + Ly * (z % Lz))) * self.spuck + n

    def nr2lattice(self, nr, _colon=None):
        cell, n = divmod(nr - 1, self.spuck)
        Lx, Ly = self.size[0], self.size[1]
        return Vec((cell % Lx, (cell // Lx) % Ly, cell // (Lx * Ly), n + 1))

    def get_species(self, site):
        return self.lattice[self.lattice2nr(site) - 1]

    def reset_site(self, site, old):
        self.lattice[self.lattice2nr(site) - 1] = self.null_species

    def replace_species(self, site, old, new):
        i = self.lattice2nr(site) - 1
        assert self.lattice[i] == old, "replace_species: found %r, expected %r at %r" % (self.lattice[i], old, site)
        self.lattice[i] = new

    # ---- base module -------------------------------------------------------------------------------------
    def avail_sites_ref(self, proc, k, plane):
        return self.avail1[proc][k] if plane == 1 else self.avail2[proc][k]

    def can_do(self, proc, site):
        return self.avail2[proc][self.lattice2nr(site)] != 0

    def add_proc(self, proc, site, rate=None):
        if proc == 0:
            return
        s = self.lattice2nr(site)
        assert self.avail2[proc][s] == 0, "add_proc: tried to add ability that is already there"
        self.nr_of_sites[proc] += 1
        n = self.nr_of_sites[proc]
        self.avail1[proc][n] = s
        self.avail2[proc][s] = n
        if self.otf:
            rm = self.rates_matrix[proc]
            rm[self.volume + 1] = rm[self.volume + 1] + rate
            rm[n] = rate

    def del_proc(self, proc, site):
        if proc == 0:
            return
        s = self.lattice2nr(site)
        memory_address = self.avail2[proc][s]
        assert memory_address != 0, "del_proc: tried to take ability from site that is not there"
        n = self.nr_of_sites[proc]
        if memory_address < n:
            self.avail1[proc][memory_address] = self.avail1[proc][n]
            self.avail1[proc][n] = 0
            if self.otf:
                rm = self.rates_matrix[proc]
                rm[self.volume + 1] = rm[self.volume + 1] - rm[memory_address]
                rm[memory_address] = rm[n]
                rm[n] = 0.0
            self.avail2[proc][self.avail1[proc][memory_address]] = memory_address
        else:
            self.avail1[proc][memory_address] = 0
            if self.otf:
                rm = self.rates_matrix[proc]
                rm[self.volume + 1] = rm[self.volume + 1] - rm[memory_address]
                rm[memory_address] = 0.0
        self.avail2[proc][s] = 0
        self.nr_of_sites[proc] = n - 1

    def update_rates_matrix(self, proc, site, rate):
        memory_address = self.avail2[proc][self.lattice2nr(site)]
        assert memory_address != 0
        rm = self.rates_matrix[proc]
        rm[self.volume + 1] = rm[self.volume + 1] + rate - rm[memory_address]
        rm[memory_address] = rate

    def increment_procstat(self, proc):
        self.procstat[proc] += 1

    def update_accum_rate(self):
        acc = self.accum_rates
        if self.otf:
            for i in range(1, self.P + 1):
                rm = self.rates_matrix[i]
                tot = 0.0
                for j in range(1, self.nr_of_sites[i] + 1):
                    tot = tot + rm[j]
                rm[self.volume + 1] = tot
                acc[i] = (acc[i - 1] + tot) if i > 1 else tot
        else:
            acc[1] = self.nr_of_sites[1] * self.rates(1)
            for i in range(2, self.P + 1):
                acc[i] = acc[i - 1] + self.nr_of_sites[i] * self.rates(i)

    @staticmethod
    def interval_search_real(arr, value):
        """arr: 1-based list (arr[0] unused) -> index (base.mpy:1234-1338)."""
        size = len(arr) - 1
        left, right = 1, size
        while True:
            mid = (right + left) >> 1
            if left >= right:
                break
            if value < arr[mid]:
                right = mid
            else:
                left = mid + 1
        if arr[mid] == 0.0:
            while not arr[mid] > 0.0:
                mid += 1
                assert mid <= size, "interval_search_real can't find available process"
        while mid != 1 and arr[mid - 1] >= arr[mid]:
            mid -= 1
        return mid

    def determine_procsite(self, ran_proc, ran_site):
        proc = self.interval_search_real(self.accum_rates, ran_proc * self.accum_rates[self.P])
        n = self.nr_of_sites[proc]
        if self.otf:
            rm = self.rates_matrix[proc]
            accp = [0.0] * (n + 1)
            accp[1] = rm[1]
            for i in range(2, n + 1):
                accp[i] = accp[i - 1] + rm[i]
            k = self.interval_search_real(accp, ran_site * accp[n])
        else:
            k = min(n, int(1 + ran_site * n))
        return proc, self.avail1[proc][k]

    def step(self):
        """One pass of do_kmc_steps' loop body -> (proc, site)."""
        ran_time, ran_proc, ran_site = self.philox_step(self.seed, self.replica, self.kmc_step)
        self.update_accum_rate()
        self.kmc_time = self.kmc_time + (-math.log(ran_time) / self.accum_rates[self.P])
        self.kmc_step += 1
        proc, site = self.determine_procsite(ran_proc, ran_site)
        self.call("run_proc_nr", [proc, site])
        return proc, site

    # ---- interpreter -------------------------------------------------------------------------------------
    def _make_function(self, name):
        def f(*args):
            return self.call(name, list(args))
        return f

    def call(self, name, args):
        if name in self.builtin:
            return self.builtin[name](*args)
        r = self.m.routine(name)
        env = dict(zip(r.params, args))
        for v, n in r.local_arrays.items():
            if v not in env:
                env[v] = Arr([0] * n)
        try:
            self._run(r.body, env)
        except _Return:
            pass
        return env.get(r.name) if r.kind == "function" else None

    def _run(self, block, env):
        g = self.g
        for st in block:
            k = st[0]
            if k == "call":
                if st[1] != "random_seed":  # the shared Philox stream replaces gfortran's generator
                    self.call(st[1], [eval(a, g, env) for a in st[2]])
            elif k == "select":
                v = eval(st[1], g, env)
                # `case default` is taken only when no other label matches, wherever it stands in the text
                # (the lat_int generator writes it in front of the last label)
                chosen = None
                for keys, body in st[2]:
                    if keys is not None and any(eval(c, g, env) == v for c in keys):
                        chosen = body
                        break
                if chosen is None:
                    chosen = next((body for keys, body in st[2] if keys is None), None)
                if chosen is not None:
                    self._run(chosen, env)
            elif k == "if":
                self._run(st[2] if eval(st[1], g, env) else st[3], env)
            elif k == "assign":
                val = eval(st[3], g, env)
                if st[2] is None:
                    env[st[1]] = val
                elif st[2] == ":":
                    arr = env[st[1]]
                    for i in range(len(arr)):
                        arr[i] = val
                else:
                    env[st[1]][eval(st[2], g, env) - 1] = val
            elif k == "return":
                raise _Return()
            elif k == "do":
                for i in range(eval(st[2], g, env), eval(st[3], g, env) + 1):
                    env[st[1]] = i
                    self._run(st[4], env)
            elif k == "decl":
                pass

    # ---- views for the comparison with the oracle ----------------------------------------------------------
    def avail_sites_array(self):
        import numpy as np
        out = np.zeros((self.P, self.volume, 2), dtype=np.int32)
        for p in range(1, self.P + 1):
            out[p - 1, :, 0] = self.avail1[p][1:self.volume + 1]
            out[p - 1, :, 1] = self.avail2[p][1:self.volume + 1]
        return out
