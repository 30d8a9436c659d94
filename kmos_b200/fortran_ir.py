"""Generated-Fortran -> rule-table IR.

kmos turns a model into Fortran (`proclist.f90`, `run_proc_*.f90`, `nli_*.f90`, `proclist_pars.f90`)
whose *statement order* defines the order of ``add_proc``/``del_proc`` calls and therefore the order of
``avail_sites`` -- which ``determine_procsite`` samples from (reference: kmos/fortran_src/base.mpy:211-302,
1075-1120).  To reproduce trajectories bit-exactly the engine must execute the very same statements in the
very same order, so instead of re-deriving the rules from the model description we read them back from the
text the reference generator wrote (generators: kmos/io/__init__.py:305-465, 1626-2057, 2219-2443,
2657-2964, 3124-3596).

The IR is plain JSON-able Python:

  off4          [dx, dy, dz, dn]      4-vector added to the routine's base coordinate
  stmt          ["replace", off4, old, new]
                ["if_can", proc, off4, [stmt...]]          guard avail_sites(proc, site, 2) /= 0
                ["del", procx, off4]
                ["add", procx, off4, ratex|None]
                ["update_rate", proc, off4, ratex]           (otf)
                ["select", off4, [[species...]|None, [stmt...]]...]   None = `case default`
                ["del_all", off4]                            strip every process from a site
                ["call", routine, off4]
                ["return", value]                            (nli functions)
                ["inc", k]                                   (gr functions: nr_vars(k) += 1, k 0-based)
  procx         int  |  ["nli", func, off4]
  ratex         ["gr", func, off4]

Process ids are 1-based and species ids 0-based exactly as in the generated constants
(kmos/fortran_src/proclist_constants.mpy:67-81).
"""
import os
import re

__all__ = ["parse_export_dir", "FortranIRError"]


class FortranIRError(Exception):
    pass


# --------------------------------------------------------------------------------------------------
# text level
# --------------------------------------------------------------------------------------------------

def _logical_lines(text):
    """Strip comments, join `&` continuations, drop blank lines."""
    out = []
    pending = ""
    for raw in text.splitlines():
        line = raw
        # strip comments (no '!' inside strings in the routines we parse; banner prints are skipped
        # because a quoted '!' never occurs there either)
        in_str = None
        cut = None
        for i, ch in enumerate(line):
            if in_str:
                if ch == in_str:
                    in_str = None
            elif ch in "\"'":
                in_str = ch
            elif ch == "!":
                cut = i
                break
        if cut is not None:
            line = line[:cut]
        line = line.strip()
        if not line:
            continue
        if pending:
            if line.startswith("&"):
                line = line[1:].lstrip()
            line = pending + " " + line
            pending = ""
        if line.endswith("&"):
            pending = line[:-1].rstrip()
            continue
        for part in _split_semicolon(line):
            out.append(part)
    if pending:
        out.append(pending)
    return out


def _split_semicolon(line):
    if ";" not in line or '"' in line or "'" in line:
        return [line]
    return [p.strip() for p in line.split(";") if p.strip()]


_CONST_RE = re.compile(
    r"^integer\(kind=iint\)\s*(?:,\s*parameter)?\s*(?:,\s*public)?\s*::\s*(\w+)\s*=\s*([-\w]+)\s*$", re.I)


def _constants(lines):
    """All `integer(kind=iint)[, parameter][, public] :: name = value` in order."""
    out = []
    for ln in lines:
        m = _CONST_RE.match(ln)
        if m:
            out.append((m.group(1), m.group(2)))
    return out


_ROUTINE_START = re.compile(r"^(?:pure\s+)?(subroutine|function)\s+(\w+)\s*\(([^)]*)\)", re.I)
_ROUTINE_START_NOARG = re.compile(r"^(?:pure\s+)?(subroutine|function)\s+(\w+)\s*$", re.I)
_ROUTINE_END = re.compile(r"^end\s+(subroutine|function)\b", re.I)


def _routines(lines):
    """name -> (kind, [args], [body lines])"""
    out = {}
    cur = None
    for ln in lines:
        if cur is None:
            m = _ROUTINE_START.match(ln) or _ROUTINE_START_NOARG.match(ln)
            if m:
                args = m.group(3) if m.lastindex and m.lastindex >= 3 else ""
                cur = (m.group(2), m.group(1).lower(), [a.strip() for a in args.split(",") if a.strip()], [])
            continue
        if _ROUTINE_END.match(ln):
            out[cur[0]] = (cur[1], cur[2], cur[3])
            cur = None
            continue
        cur[3].append(ln)
    return out


# --------------------------------------------------------------------------------------------------
# expression level
# --------------------------------------------------------------------------------------------------

class _Env(object):
    """Name resolution for one export: species / process / site / layer constants."""

    def __init__(self):
        self.species = {}
        self.procs = {}
        self.sites = {}     # `<layer>_<site>` -> 1-based index in cell
        self.layers = {}
        self.misc = {}
        self.userpar = {}
        self.chempots = {}

    def _lookup(self, name):
        for d in (self.sites, self.layers, self.misc):
            if name in d:
                return d[name]
        low = name.lower()
        for d in (self.sites, self.layers, self.misc):
            for k, v in d.items():
                if k.lower() == low:
                    return v
        raise FortranIRError("unknown integer constant %r" % name)

    def int_expr(self, expr):
        """Evaluate a small integer expression like `ruo2_cus - ruo2_bridge` or `+0`."""
        expr = expr.strip()
        toks = re.findall(r"\w+|[-+()]", expr)
        if "".join(toks) != expr.replace(" ", ""):
            raise FortranIRError("cannot evaluate %r" % expr)
        py = []
        for t in toks:
            if re.match(r"^\d+$", t) or t in "+-()":
                py.append(t)
            else:
                py.append(str(self._lookup(t)))
        return int(eval(" ".join(py), {"__builtins__": {}}))

    def species_id(self, name):
        name = name.strip()
        if name in self.species:
            return self.species[name]
        for k, v in self.species.items():
            if k.lower() == name.lower():
                return v
        if name == "null_species":
            return -1
        if name == "default_species":
            return self.misc["default_species"]
        raise FortranIRError("unknown species %r" % name)

    def proc_id(self, name):
        name = name.strip()
        if name in self.procs:
            return self.procs[name]
        for k, v in self.procs.items():
            if k.lower() == name.lower():
                return v
        raise FortranIRError("unknown process %r" % name)


_VEC_RE = re.compile(r"\(/(.*?)/\)")


def _split_top(s, sep=","):
    """Split on `sep` at parenthesis depth 0 (treating `(/ ... /)` like parentheses)."""
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def _coord(expr, env, variables):
    """`site`, `lsite + (/0, 1, 0, 0/)`, `cell + (/+0, -1, +0, ruo2_cus/)` -> off4 relative to the base."""
    expr = expr.strip()
    m = re.match(r"^(\w+)\s*(?:\+\s*\(/(.*)/\))?$", expr)
    if not m:
        raise FortranIRError("cannot parse coordinate %r" % expr)
    base = m.group(1)
    if base not in variables:
        raise FortranIRError("unknown coordinate variable %r in %r" % (base, expr))
    off = list(variables[base])
    if m.group(2) is not None:
        comps = _split_top(m.group(2))
        if len(comps) != 4:
            raise FortranIRError("expected 4 components in %r" % expr)
        for i, c in enumerate(comps):
            off[i] += env.int_expr(c)
    return off


def _find_call_args(s, start):
    """s[start] == '(' -> (inner text, index after the matching ')')"""
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return s[start + 1:i], i + 1
    raise FortranIRError("unbalanced parentheses in %r" % s)


def _procx(expr, env, variables):
    expr = expr.strip()
    m = re.match(r"^(nli_\w+)\s*\(", expr)
    if m:
        inner, _ = _find_call_args(expr, m.end() - 1)
        return ["nli", m.group(1), _coord(inner, env, variables)]
    if expr == "proc_nr":
        raise FortranIRError("loop variable process outside del_all idiom")
    return env.proc_id(expr)


def _ratex(expr, env, variables):
    expr = expr.strip()
    m = re.match(r"^(gr_\w+)\s*\(", expr)
    if not m:
        raise FortranIRError("cannot parse rate expression %r" % expr)
    inner, _ = _find_call_args(expr, m.end() - 1)
    return ["gr", m.group(1), _coord(inner, env, variables)]


def _lattice2nr_offset(args, variables, env=None):
    """`site(1) + (0), site(2) + (-1), site(3) + (0), site(4) + (0)` -> off4."""
    comps = _split_top(args)
    if len(comps) != 4:
        raise FortranIRError("lattice2nr with %d args" % len(comps))
    off = None
    for i, c in enumerate(comps):
        m = re.match(r"^(\w+)\((\d)\)\s*(?:\+\s*\((.*)\))?$", c.strip())
        if not m or int(m.group(2)) != i + 1:
            raise FortranIRError("cannot parse lattice2nr arg %r" % c)
        if off is None:
            off = list(variables[m.group(1)])
        if m.group(3):
            off[i] += env.int_expr(m.group(3)) if env is not None else int(m.group(3))
    return off


# --------------------------------------------------------------------------------------------------
# statement level
# --------------------------------------------------------------------------------------------------

_IGNORED = re.compile(
    r"^(integer|real|logical|character|implicit|use\b|print|return$|nr_vars\(:\)\s*=\s*0)", re.I)


class _BlockParser(object):
    def __init__(self, lines, env, variables, fname=None):
        self.lines = lines
        self.i = 0
        self.env = env
        self.variables = dict(variables)
        self.fname = fname

    def peek(self):
        return self.lines[self.i] if self.i < len(self.lines) else None

    def parse_block(self, terminators=()):
        stmts = []
        while self.i < len(self.lines):
            ln = self.lines[self.i]
            low = ln.lower()
            if any(low.startswith(t) for t in terminators):
                return stmts
            self.i += 1
            st = self.parse_stmt(ln)
            if st is not None:
                if isinstance(st, list) and st and isinstance(st[0], list):
                    stmts.extend(st)
                else:
                    stmts.append(st)
        if terminators:
            raise FortranIRError("missing terminator %r" % (terminators,))
        return stmts

    def parse_stmt(self, ln):
        env, variables = self.env, self.variables
        low = ln.lower()
        if _IGNORED.match(ln):
            return None
        # ---- calls
        m = re.match(r"^call\s+(\w+)\s*\((.*)\)\s*$", ln, re.I)
        if m:
            name, args = m.group(1), _split_top(m.group(2))
            lname = name.lower()
            if lname == "replace_species":
                return ["replace", _coord(args[0], env, variables), env.species_id(args[1]),
                        env.species_id(args[2])]
            if lname == "del_proc":
                return ["del", _procx(args[0], env, variables), _coord(args[1], env, variables)]
            if lname == "add_proc":
                rate = _ratex(args[2], env, variables) if len(args) > 2 else None
                return ["add", _procx(args[0], env, variables), _coord(args[1], env, variables), rate]
            if lname == "update_rates_matrix":
                return ["update_rate", env.proc_id(args[0]), _coord(args[1], env, variables),
                        _ratex(args[2], env, variables)]
            if lname == "increment_procstat":
                return None
            if lname.startswith(("create_", "annihilate_")) and len(args) == 2:
                # multi-lattice models (kmos/io/__init__.py:2445-2560): create_<site>(site, species) /
                # annihilate_<site>(site, species) -- one instance per species actually passed
                return ["call", "%s@%d" % (name, env.species_id(args[1])), _coord(args[0], env, variables)]
            return ["call", name, _coord(args[0], env, variables)]
        m = re.match(r"^call\s+(\w+)\s*$", ln, re.I)
        if m:
            return None
        # ---- select case
        m = re.match(r"^select\s+case\s*\(\s*get_species\s*\((.*)\)\s*\)\s*$", ln, re.I)
        if m:
            off = _coord(m.group(1), env, variables)
            cases = []
            while True:
                nxt = self.peek()
                if nxt is None:
                    raise FortranIRError("unterminated select")
                nlow = nxt.lower()
                if re.match(r"^end\s*select", nlow):
                    self.i += 1
                    break
                mc = re.match(r"^case\s*\((.*)\)\s*$", nxt, re.I)
                md = re.match(r"^case\s+default\s*$", nxt, re.I)
                if not (mc or md):
                    raise FortranIRError("expected case, got %r" % nxt)
                self.i += 1
                key = None if md else [env.species_id(s) for s in _split_top(mc.group(1))]
                body = self.parse_block(terminators=("case", "end select", "endselect"))
                cases.append([key, body])
            return ["select", off, cases]
        # ---- guarded if
        m = re.match(r"^if\s*\((.*)\)\s*then\s*$", ln, re.I)
        if m:
            cond = m.group(1).strip()
            body = self.parse_block(terminators=("endif", "end if"))
            self.i += 1  # consume endif
            mc = re.match(r"^avail_sites\s*\(\s*(\w+)\s*,\s*lattice2nr\s*\((.*)\)\s*,\s*2\s*\)\s*\.ne\.\s*0$",
                          cond, re.I)
            if mc:
                pname = mc.group(1)
                off = _lattice2nr_offset(mc.group(2), variables, env)
                if pname == "proc_nr":
                    return ("__loop_guard__", off, body)
                return ["if_can", env.proc_id(pname), off, body]
            mc = re.match(r"^can_do\s*\((.*)\)$", cond, re.I)
            if mc:
                a = _split_top(mc.group(1))
                return ["if_can", env.proc_id(a[0]), _coord(a[1], env, variables), body]
            raise FortranIRError("unsupported if condition %r" % cond)
        # ---- the `do proc_nr = 1, nr_of_proc` strip-all idiom of touchup_cell
        m = re.match(r"^do\s+proc_nr\s*=\s*1\s*,\s*nr_of_proc\s*$", ln, re.I)
        if m:
            # body is parsed with the loop variable unresolved: handle textually
            body_lines = []
            while self.peek() is not None and not re.match(r"^end\s*do", self.peek(), re.I):
                body_lines.append(self.peek())
                self.i += 1
            self.i += 1
            joined = " ".join(body_lines)
            mc = re.search(r"lattice2nr\s*\((.*?)\)\s*,\s*2\s*\)", joined)
            if not mc or "del_proc(proc_nr" not in joined.replace(" ", ""):
                raise FortranIRError("unrecognised do-loop %r" % joined)
            return ["del_all", _lattice2nr_offset(mc.group(1), variables, env)]
        # ---- assignments
        m = re.match(r"^(\w+)\s*=\s*(.*)$", ln)
        if m:
            lhs, rhs = m.group(1), m.group(2).strip()
            if self.fname and lhs.lower() == self.fname.lower():
                if re.match(r"^rate_\w+\s*\(", rhs):
                    return None  # gr_x = rate_x(nr_vars)
                if rhs == "0":
                    return ["return", 0]
                return ["return", env.proc_id(rhs)]
            if lhs in ("site", "cell", "lsite"):
                mm = re.match(r"^nr2lattice\s*\(\s*\w+\s*,\s*:\s*\)\s*(?:\+\s*\(/(.*)/\))?$", rhs)
                if mm:
                    off = [0, 0, 0, 0]
                    if mm.group(1):
                        off = [env.int_expr(c) for c in _split_top(mm.group(1))]
                    self.variables[lhs] = off
                    return None
                self.variables[lhs] = _coord(rhs, env, self.variables)
                return None
        m = re.match(r"^nr_vars\((\d+)\)\s*=\s*nr_vars\((\d+)\)\s*\+\s*1$", ln)
        if m:
            return ["inc", int(m.group(1)) - 1]
        raise FortranIRError("unsupported statement %r" % ln)


def _parse_routine(body, env, variables, fname=None):
    bp = _BlockParser(body, env, variables, fname)
    return bp.parse_block()


# --------------------------------------------------------------------------------------------------
# file level
# --------------------------------------------------------------------------------------------------

def _read(path):
    with open(path) as f:
        return f.read()


def _site_of_routine(name, env):
    """`put_co_ruo2_bridge` -> site index of `ruo2_bridge` (longest matching suffix)."""
    name = name.partition("@")[0]
    best = None
    for sname, idx in env.sites.items():
        if name.lower().endswith("_" + sname.lower()):
            if best is None or len(sname) > len(best[0]):
                best = (sname, idx)
    return best[1] if best else None


def parse_export_dir(path, backend=None):
    """Parse a kmos export directory (the `src/` the reference would hand to f2py) into the IR."""
    files = sorted(os.listdir(path))
    if backend is None:
        if "proclist_pars.f90" in files:
            backend = "otf"
        elif any(f.startswith("nli_") for f in files):
            backend = "lat_int"
        else:
            backend = "local_smart"
    env = _Env()

    lat_lines = _logical_lines(_read(os.path.join(path, "lattice.f90")))
    proclist_lines = _logical_lines(_read(os.path.join(path, "proclist.f90")))
    const_lines = proclist_lines
    if "proclist_constants.f90" in files:
        const_lines = _logical_lines(_read(os.path.join(path, "proclist_constants.f90"))) + proclist_lines

    # ---- lattice constants (kmos/fortran_src/lattice.mpy:60-132)
    lat_consts = _constants(lat_lines)
    lat = dict(lat_consts)
    nr_of_layers = int(lat["nr_of_layers"])
    model_dimension = int(lat["model_dimension"])
    spuck = int(lat["spuck"])
    names = [n for n, _ in lat_consts]
    i_md = names.index("model_dimension")
    layer_names = names[i_md + 1:i_md + 1 + nr_of_layers]
    for n in layer_names:
        env.layers[n] = int(lat[n])
    env.misc["default_layer"] = env.layers[lat["default_layer"]] if lat["default_layer"] in env.layers \
        else int(lat["default_layer"])
    site_names = []
    for n, v in lat_consts:
        if n in ("nr_of_layers", "model_dimension", "spuck", "default_layer", "substrate_layer") \
                or n in env.layers:
            continue
        env.sites[n] = int(v)
        site_names.append(n)
    if sorted(env.sites.values()) != list(range(1, spuck + 1)):
        raise FortranIRError("site constants %r do not cover 1..spuck=%d" % (env.sites, spuck))
    site_names.sort(key=lambda n: env.sites[n])

    # ---- species / process constants
    consts = _constants(const_lines)
    cdict = dict(consts)
    nr_of_species = int(cdict["nr_of_species"])
    nr_of_proc = int(cdict["nr_of_proc"])
    cnames = [n for n, _ in consts]
    i_sp = cnames.index("nr_of_species")
    species_names = cnames[i_sp + 1:i_sp + 1 + nr_of_species]
    for n in species_names:
        env.species[n] = int(cdict[n])
    if sorted(env.species.values()) != list(range(nr_of_species)):
        raise FortranIRError("species ids are not 0..n-1: %r" % env.species)
    dflt = cdict.get("default_species")
    env.misc["default_species"] = env.species[dflt] if dflt in env.species else int(dflt)
    rest = [(n, v) for n, v in consts[i_sp + 1 + nr_of_species:]
            if n not in ("default_species", "representation_length", "seed_size", "nr_of_proc")]
    proc_names = []
    for n, v in rest:
        if len(proc_names) == nr_of_proc:
            break
        if int(v) == len(proc_names) + 1:
            env.procs[n] = int(v)
            proc_names.append(n)
    if len(proc_names) != nr_of_proc:
        raise FortranIRError("found %d process constants, expected %d" % (len(proc_names), nr_of_proc))

    ir = {
        "ir_version": 1,
        "backend": backend,
        "model_dimension": model_dimension,
        "spuck": spuck,
        "species": sorted(env.species, key=lambda n: env.species[n]),
        "default_species": env.misc["default_species"],
        # multi-lattice models declare a species called null_species and hand it to base.set_null_species
        # (proclist_generic_subroutines.mpy:223): sites outside the initial layer hold this id, not -1
        "null_species": next((v for k, v in env.species.items() if k.lower() == "null_species"), -1),
        "layers": sorted(env.layers, key=lambda n: env.layers[n]),
        "default_layer": env.misc["default_layer"],
        "sites": site_names,
        "procs": proc_names,
        "routines": {},
        "nli": {},
        "gr": {},
    }

    routines = _routines(proclist_lines)
    extra = {}
    for f in files:
        if re.match(r"^(run_proc|nli)_\d+\.f90$", f):
            extra.update(_routines(_logical_lines(_read(os.path.join(path, f)))))

    # ---- initialize_state (kmos/fortran_src/proclist_generic_subroutines.mpy:236-304)
    ir["init"] = _parse_initialize_state(routines["initialize_state"][2], env)

    # ---- run_proc_nr
    kind, args, body = routines["run_proc_nr"]
    ir["run_proc"] = _parse_run_proc_nr(body, env, nr_of_proc)

    # ---- the routines run_proc_nr / initialize_state call
    def need(name):
        if name in ir["routines"]:
            return
        base, _, sp = name.partition("@")
        src = routines.get(base) or extra.get(base)
        if src is None:
            raise FortranIRError("routine %r not found" % base)
        argname = src[1][0] if src[1] else "site"
        if sp:  # create_/annihilate_ instance: bind the routine's species argument to the constant passed
            spname = src[1][1] if len(src[1]) > 1 else "species"
            if spname in env.species:
                raise FortranIRError("species argument %r shadows a species constant" % spname)
            env.species[spname] = int(sp)
            try:
                stmts = _parse_routine(src[2], env, {argname: [0, 0, 0, 0]})
            finally:
                del env.species[spname]
        else:
            stmts = _parse_routine(src[2], env, {argname: [0, 0, 0, 0]})
        ir["routines"][name] = stmts
        for callee in _callees(stmts):
            need(callee)

    for calls in ir["run_proc"]:
        for c in calls:
            need(c[1])
    for layer in ir["init"]["layers"].values():
        for r, _off in layer["touchups"]:
            need(r)

    # ---- nli functions (lat_int)
    for name, (kind, args, body) in extra.items():
        if kind == "function" and name.lower().startswith("nli_"):
            ir["nli"][name] = _parse_routine(body, env, {args[0]: [0, 0, 0, 0]}, fname=name)

    # ---- otf: gr_/rate_ functions and parameter tables
    if backend == "otf":
        _parse_pars(os.path.join(path, "proclist_pars.f90"), env, ir)

    if backend == "local_smart":
        ir["routine_site"] = {}
        for name in ir["routines"]:
            s = _site_of_routine(name, env)
            if s is not None:
                ir["routine_site"][name] = s
    return ir


def _callees(stmts):
    for st in stmts:
        if st[0] == "call":
            yield st[1]
        elif st[0] == "select":
            for _k, body in st[2]:
                for c in _callees(body):
                    yield c
        elif st[0] == "if_can":
            for c in _callees(st[3]):
                yield c


def _parse_run_proc_nr(body, env, nr_of_proc):
    variables = {}
    calls = [None] * nr_of_proc
    cur = None
    for ln in body:
        low = ln.lower()
        if _IGNORED.match(ln) or low.startswith("call increment_procstat"):
            continue
        m = re.match(r"^(\w+)\s*=\s*nr2lattice\s*\(\s*\w+\s*,\s*:\s*\)\s*(?:\+\s*\(/(.*)/\))?$", ln)
        if m:
            off = [0, 0, 0, 0]
            if m.group(2):
                off = [env.int_expr(c) for c in _split_top(m.group(2))]
            variables[m.group(1)] = off
            continue
        if low.startswith("select case"):
            continue
        if re.match(r"^end\s*select", low):
            cur = None
            continue
        m = re.match(r"^case\s*\(([\w\s,]+)\)$", ln, re.I)
        if m:
            # lat_int groups share one run_proc routine: `case(p_0, p_1, ...)`
            cur = [env.proc_id(n) - 1 for n in _split_top(m.group(1))]
            shared = []
            for c in cur:
                calls[c] = shared
            continue
        if re.match(r"^case\s+default", low):
            cur = None
            continue
        if low in ("stop",):
            continue
        m = re.match(r"^call\s+(\w+)\s*\((.*)\)\s*$", ln, re.I)
        if m and cur is not None:
            args = _split_top(m.group(2))
            cname = m.group(1)
            if cname.lower().startswith(("create_", "annihilate_")) and len(args) == 2:
                cname = "%s@%d" % (cname, env.species_id(args[1]))
            calls[cur[0]].append(["call", cname, _coord(args[0], env, variables)])
            continue
        if cur is None:
            continue
        raise FortranIRError("unsupported line in run_proc_nr: %r" % ln)
    for i, c in enumerate(calls):
        if c is None:
            raise FortranIRError("process %d has no case in run_proc_nr" % (i + 1))
    return calls


def _parse_initialize_state(body, env):
    layers = {}
    cur = None
    cell_touchup = False
    for ln in body:
        m = re.match(r"^case\s*\(\s*(\w+)\s*\)$", ln, re.I)
        if m and m.group(1) in env.layers:
            cur = layers.setdefault(str(env.layers[m.group(1)]), {"defaults": [], "touchups": []})
            continue
        m = re.match(r"^call\s+replace_species\s*\(\s*\(/\s*i\s*,\s*j\s*,\s*k\s*,\s*(\w+)\s*/\)\s*,\s*null_species\s*,"
                     r"\s*(\w+)\s*\)$", ln, re.I)
        if m and cur is not None:
            cur["defaults"].append([env.int_expr(m.group(1)), env.species_id(m.group(2))])
            continue
        m = re.match(r"^call\s+(touchup_\w+)\s*\(\s*\(/\s*i\s*,\s*j\s*,\s*k\s*,\s*(\w+)\s*/\)\s*\)$", ln, re.I)
        if m:
            off = [0, 0, 0, env.int_expr(m.group(2))]
            if m.group(1).lower() == "touchup_cell":
                cell_touchup = True
                for lay in layers.values():
                    lay["touchups"].append([m.group(1), off])
            elif cur is not None:
                cur["touchups"].append([m.group(1), off])
            continue
    if not layers:
        raise FortranIRError("initialize_state: no layer cases found")
    return {"layers": layers, "cell_touchup": cell_touchup}


def _parse_pars(path, env, ir):
    lines = _logical_lines(_read(path))
    userpar, chempots, fconsts = [], [], {}
    for ln in lines:
        m = re.match(r"^integer\(kind=iint\)\s*,\s*public\s*::\s*(\w+)\s*=\s*(\d+)$", ln)
        if m:
            userpar.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"^real\(kind=rdouble\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*(\S+)$", ln)
        if m:
            fconsts[m.group(1)] = m.group(2)
    # user parameters come first, then chemical potentials (kmos/io/__init__.py:2657-2760)
    n_userpar = n_chempots = 0
    for ln in lines:
        m = re.match(r"^real\(kind=rdouble\)\s*,\s*public\s*,\s*dimension\((\d+)\)\s*::\s*(\w+)$", ln)
        if m and m.group(2) == "userpar":
            n_userpar = int(m.group(1))
        if m and m.group(2) == "chempots":
            n_chempots = int(m.group(1))
    ir["userpar"] = [n for n, _ in userpar[:n_userpar]]
    ir["chempots"] = [n for n, _ in userpar[n_userpar:n_userpar + n_chempots]]
    ir["pars_constants"] = fconsts
    # byst_<proc>: the names of gr_<proc>'s nr_vars counters, in order (kmos/io/__init__.py:2842-2848); the
    # front-end reads them back to label rate_<proc>'s argument (kmos/run/__init__.py:1868-1899)
    ir["byst"] = {}
    for ln in lines:
        m = re.match(r'^character\(len=\d+\)\s*,\s*parameter\s*,\s*public\s*::\s*byst_(\w+)\s*=\s*"(.*)"$', ln)
        if m:
            ir["byst"][m.group(1)] = " ".join(m.group(2).split())
    routines = _routines(lines)
    ir["rate_expr"] = {}
    for name, (kind, args, body) in routines.items():
        if kind != "function":
            continue
        if name.startswith("gr_"):
            nvars = 0
            for ln in body:
                m = re.match(r"^integer\(kind=iint\)\s*,\s*dimension\((\d+)\)\s*::\s*nr_vars$", ln)
                if m:
                    nvars = int(m.group(1))
            stmts = _parse_routine(body, env, {args[0]: [0, 0, 0, 0]}, fname=name)
            ir["gr"][name] = {"nvars": nvars, "body": stmts}
        elif name.startswith("rate_"):
            expr = None
            for ln in body:
                m = re.match(r"^%s\s*=\s*(.*)$" % re.escape(name), ln)
                if m:
                    expr = m.group(1).strip()
            ir["rate_expr"][name[len("rate_"):]] = expr
