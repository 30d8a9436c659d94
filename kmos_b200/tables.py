"""Rule-table IR (kmos_b200.fortran_ir) -> binary model blob.

The blob is what crosses the C-ABI (`kmos_b200_model_create`, include/kmos_b200.h) and what the CPU
oracle loads.  It is a flat little-endian int32 array:

    [0]  magic 0x4B423230 ("KB20")     [1] version          [2] backend (0 local_smart,1 lat_int,2 otf)
    [3]  n_species   [4] n_proc   [5] spuck   [6] model_dimension   [7] default_species | (null_species id + 1) << 16
    [8]  n_layers    [9] default_layer   [10] n_routines   [11] n_gr   [12] lut_total
    [13] n_sections  then n_sections x (id, offset_words, length_words)

Sections
    SEC_ROUTINES   n_routines x (code_offset, code_len)   offsets into SEC_CODE
    SEC_CODE       statement byte-code (see OP_*): a 1:1 encoding of the generated Fortran statements,
                   in textual order -- the order that defines avail_sites (base.mpy:211-302)
    SEC_RUNPROC    n_proc routine ids: run_proc_nr's `select case(proc)` (io/__init__.py:305-465)
    SEC_INIT       n_layers x (defaults_routine, touchup_routine): initialize_state's two loop bodies
                   (proclist_generic_subroutines.mpy:261-299)
    SEC_GR         otf: n_gr x (routine, proc, nvars, lut_offset, radix[MAX_VARS])
    SEC_PROCSITE   n_proc x site type (1-based) the process is registered on (0: several -> unsupported)
    SEC_DEVICE     compiled per-event lane tables for the CUDA engine (kmos_b200.devtables)
    SEC_DEVICE_HBM local_smart only: the same events in textual order for the HBM-resident warp kernel

Coordinates are 4-vectors [dx,dy,dz,dn] added to the routine's base coordinate, exactly like the
`site + (/dx,dy,dz,dn/)` expressions of the generated code (lattice.mpy:343-491).
"""
import numpy as np

MAGIC = 0x4B423230
VERSION = 4

BACKENDS = {"local_smart": 0, "lat_int": 1, "otf": 2}

SEC_ROUTINES, SEC_CODE, SEC_RUNPROC, SEC_INIT, SEC_GR, SEC_PROCSITE, SEC_DEVICE, SEC_DEVICE_HBM = range(1, 9)

(OP_REPLACE, OP_IF_CAN, OP_DEL, OP_ADD, OP_DEL_NLI, OP_ADD_NLI, OP_ADD_RATE, OP_UPD_RATE, OP_SELECT,
 OP_CASE, OP_DEL_ALL, OP_CALL, OP_RETURN, OP_INC, OP_JUMP) = range(1, 16)

MAX_VARS = 8
CASE_DEFAULT = -1  # mask with every bit set


class TableError(Exception):
    pass


class _Assembler(object):
    def __init__(self, ir):
        self.ir = ir
        self.routine_ids = {}
        self.routine_code = []  # list of word lists
        self.gr_ids = {}

    def routine_id(self, name):
        if name in self.routine_ids:
            return self.routine_ids[name]
        ir = self.ir
        if name in ir["routines"]:
            stmts = ir["routines"][name]
        elif name in ir["nli"]:
            stmts = ir["nli"][name]
        elif name in ir["gr"]:
            stmts = ir["gr"][name]["body"]
        else:
            raise TableError("unknown routine %r" % name)
        rid = len(self.routine_code)
        self.routine_ids[name] = rid
        self.routine_code.append(None)  # reserve (recursion-safe)
        self.routine_code[rid] = self.block(stmts)
        return rid

    def anon_routine(self, name, stmts):
        rid = len(self.routine_code)
        self.routine_ids[name] = rid
        self.routine_code.append(None)
        self.routine_code[rid] = self.block(stmts)
        return rid

    def gr_id(self, name):
        if name not in self.gr_ids:
            self.gr_ids[name] = len(self.gr_ids)
            self.routine_id(name)
        return self.gr_ids[name]

    def block(self, stmts):
        out = []
        for st in stmts:
            out.extend(self.stmt(st))
        return out

    def stmt(self, st):
        k = st[0]
        if k == "replace":
            return [OP_REPLACE] + st[1] + [st[2], st[3]]
        if k == "if_can":
            body = self.block(st[3])
            return [OP_IF_CAN, st[1]] + st[2] + [len(body)] + body
        if k == "del":
            if isinstance(st[1], list):
                return [OP_DEL_NLI, self.routine_id(st[1][1])] + st[1][2] + st[2]
            return [OP_DEL, st[1]] + st[2]
        if k == "add":
            if isinstance(st[1], list):
                if st[3] is not None:
                    raise TableError("add with nli process and rate")
                return [OP_ADD_NLI, self.routine_id(st[1][1])] + st[1][2] + st[2]
            if st[3] is not None:
                return [OP_ADD_RATE, st[1]] + st[2] + [self.gr_id(st[3][1])] + st[3][2]
            return [OP_ADD, st[1]] + st[2]
        if k == "update_rate":
            return [OP_UPD_RATE, st[1]] + st[2] + [self.gr_id(st[3][1])] + st[3][2]
        if k == "select":
            cases = []
            for key, body in st[2]:
                b = self.block(body)
                if key is None:
                    mask = CASE_DEFAULT
                else:
                    mask = 0
                    for s in key:
                        if s < 0:
                            raise TableError("case(null_species) not supported")
                        mask |= 1 << s
                cases.append([mask, b])
            # `case default` is evaluated last, wherever it was written
            cases.sort(key=lambda c: c[0] == CASE_DEFAULT)
            # every case body ends in OP_JUMP <words to skip to the end of the select>, so that a
            # non-recursive interpreter can fall out of the taken case (kmos_b200/csrc/kb_interp.h)
            flat = []
            remaining = sum(3 + len(b) + 2 for _m, b in cases)
            for mask, b in cases:
                remaining -= 3 + len(b) + 2
                flat += [OP_CASE, mask, len(b) + 2] + b + [OP_JUMP, remaining]
            return [OP_SELECT] + st[1] + [len(cases), len(flat)] + flat
        if k == "del_all":
            return [OP_DEL_ALL] + st[1]
        if k == "call":
            return [OP_CALL, self.routine_id(st[1])] + st[2]
        if k == "return":
            return [OP_RETURN, st[1]]
        if k == "inc":
            return [OP_INC, st[1]]
        raise TableError("unknown statement %r" % (st,))


def _gr_radices(stmts, nvars):
    """Upper bound (+1) of each nr_vars(k): how many `select` statements can increment it."""
    counts = [0] * max(nvars, 1)

    def walk(block):
        for st in block:
            if st[0] == "select":
                per_case = []
                for _key, body in st[2]:
                    c = [0] * len(counts)
                    for b in body:
                        if b[0] == "inc":
                            c[b[1]] += 1
                        elif b[0] == "select":
                            raise TableError("nested select in gr function")
                    per_case.append(c)
                for i in range(len(counts)):
                    counts[i] += max(c[i] for c in per_case) if per_case else 0
            elif st[0] == "inc":
                counts[st[1]] += 1
    walk(stmts)
    return [c + 1 for c in counts[:nvars]]


def proc_site_masks(ir):
    """For every process the set of site types (bit n-1) it is ever registered on."""
    nproc = len(ir["procs"])
    masks = [0] * nproc
    if ir["backend"] != "local_smart":
        return [1] * nproc  # lat_int / otf register every process on site 1 of the cell
    rsite = ir["routine_site"]

    def walk(block, base_n):
        for st in block:
            if st[0] == "add" and not isinstance(st[1], list):  # registered only where it is added
                masks[st[1] - 1] |= 1 << (base_n + st[2][3] - 1)
            elif st[0] == "if_can":
                walk(st[3], base_n)
            elif st[0] == "select":
                for _k, body in st[2]:
                    walk(body, base_n)
    for name, stmts in ir["routines"].items():
        if name in rsite:
            walk(stmts, rsite[name])
    return masks


def proc_anchor_types(ir):
    """1-based site type each process is registered on; 0 if it is registered on several (unsupported by
    the per-cell avail_sites layout of the engine)."""
    out = []
    for mask in proc_site_masks(ir):
        out.append(mask.bit_length() if mask and not (mask & (mask - 1)) else 0)
    return out


def build_blob(ir, with_device=True):
    """Assemble the int32 blob for a parsed model.  Returns (np.int32 array, info dict)."""
    asm = _Assembler(ir)
    nproc = len(ir["procs"])
    # run_proc_nr: one anonymous routine per process holding its call list
    runproc = []
    seen = {}
    for p, calls in enumerate(ir["run_proc"]):
        key = repr(calls)
        if key not in seen:
            seen[key] = asm.anon_routine("__run_proc_%d" % (p + 1), calls)
        runproc.append(seen[key])
    # initialize_state
    nlayers = len(ir["layers"])
    init = []
    for layer in range(nlayers):
        lay = ir["init"]["layers"].get(str(layer))
        if lay is None:
            init += [-1, -1]
            continue
        defaults = [["replace", [0, 0, 0, n], ir.get("null_species", -1), sp] for n, sp in lay["defaults"]]
        touch = [["call", r, off] for r, off in lay["touchups"]]
        init += [asm.anon_routine("__init_defaults_%d" % layer, defaults),
                 asm.anon_routine("__init_touchup_%d" % layer, touch)]
    # make sure every nli/gr is assembled (they may be unreachable from run_proc in tiny models)
    for name in sorted(ir["nli"]):
        asm.routine_id(name)
    gr_words = []
    lut_total = 0
    gr_info = {}
    if ir["backend"] == "otf":
        for name in sorted(ir["gr"]):
            asm.gr_id(name)
        by_id = sorted(asm.gr_ids, key=lambda n: asm.gr_ids[n])
        proc_lower = {p.lower(): i + 1 for i, p in enumerate(ir["procs"])}
        for name in by_id:
            g = ir["gr"][name]
            nvars = g["nvars"]
            if nvars > MAX_VARS:
                raise TableError("%s: %d bystander counters > MAX_VARS" % (name, nvars))
            radix = _gr_radices(g["body"], nvars)
            size = int(np.prod(radix)) if radix else 1
            proc = proc_lower[name[len("gr_"):].lower()]
            gr_words += [asm.routine_ids[name], proc, nvars, lut_total] + radix + [1] * (MAX_VARS - nvars)
            gr_info[name] = {"proc": proc, "nvars": nvars, "radix": radix, "lut_offset": lut_total,
                             "size": size}
            lut_total += size

    otf_dev = None
    if with_device and ir["backend"] == "otf":   # assembles the events' tail routines: before the code is laid out
        from . import devtables
        otf_dev = devtables.compile_otf_tables(ir, asm)

    code = []
    routines = []
    for words in asm.routine_code:
        routines += [len(code), len(words)]
        code += words

    sections = [
        (SEC_ROUTINES, routines),
        (SEC_CODE, code),
        (SEC_RUNPROC, runproc),
        (SEC_INIT, init),
        (SEC_GR, gr_words),
        (SEC_PROCSITE, proc_anchor_types(ir)),
    ]
    info = {"routine_ids": dict(asm.routine_ids), "gr": gr_info, "lut_total": lut_total,
            "n_routines": len(asm.routine_code)}
    if with_device:
        from . import devtables
        dev_words, dev_info = otf_dev if otf_dev is not None else devtables.compile_device_tables(ir, asm)
        sections.append((SEC_DEVICE, dev_words))
        info["device"] = dev_info
        if ir["backend"] == "local_smart":
            hbm_words, hbm_info = devtables.compile_hbm_tables(ir)
            sections.append((SEC_DEVICE_HBM, hbm_words))
            info["device_hbm"] = hbm_info

    header = [MAGIC, VERSION, BACKENDS[ir["backend"]], len(ir["species"]), nproc, ir["spuck"],
              ir["model_dimension"], ir["default_species"] | ((ir.get("null_species", -1) + 1) << 16), nlayers,
              ir["default_layer"],
              len(asm.routine_code), len(asm.gr_ids), lut_total, len(sections)]
    dir_len = 3 * len(sections)
    off = len(header) + dir_len
    directory = []
    body = []
    for sid, words in sections:
        directory += [sid, off, len(words)]
        body += list(words)
        off += len(words)
    blob = np.asarray(header + directory + body, dtype=np.int64)
    if np.any(blob > 2**31 - 1) or np.any(blob < -2**31):
        raise TableError("blob word out of int32 range")
    return blob.astype(np.int32), info


def load_ir(path):
    import json
    with open(path) as f:
        return json.load(f)
