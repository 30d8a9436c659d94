"""Exporter back end: emit a model's proclist as specialised CUDA (``proclist_<model>.cu``).

kmos writes ``run_proc_nr`` and the ``put_``/``take_`` routines as model-specific straight-line Fortran
(kmos/io/__init__.py:305-465 write_proclist_run_proc_nr_smart, :2219-2409 write_proclist_put_take,
:2568-2655 _write_optimal_iftree); ``export_source`` (:3884-3974, hook point :3958-3973) is where the files
are written.  This module is the CUDA twin of that step for the local_smart backend: from the same rule IR the
Fortran was parsed into (kmos_b200.fortran_ir -- so statement order, and with it the order of ``avail_sites``,
is the Fortran's) it emits

  * one ``case`` per process: the site read, the event's ``replace_species`` calls with offsets/species as
    immediates, then its guarded ``del_proc`` / if-tree ``add_proc`` calls as unrolled *rounds* whose shape
    (dels only / adds only / mixed, number of dynamic probes, number of ops, operand addresses) is fixed at
    compile time; event dispatch is a ``switch``;
  * the static operand descriptions the host turns into the geometry-specialised operand table
    (``kb_gen_fill_tables`` in csrc/kb_gen.cuh), and the model constants (process count, sites per cell,
    neighbour offsets) the step skeleton is instantiated with.

The result is compiled with nvcc for sm_100a into ``proclist_<model>_<hash>.so`` (cached by content hash) and
attached to a batch with ``kmos_b200_batch_attach_proclist``; the table interpreter (kb_smem.cuh) stays the
path for models this generator declines (``Unsupported``).

    python -m kmos_b200.codegen <model_tables.json | export_dir> [-o out_dir] [--style unrolled|compact] [--build]
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

from . import devtables, tables
from .devtables import KIND_ADD, Unsupported

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CACHE = os.path.join(HERE, "_proclist_cache")
GEN_VERSION = 3
MAX_COND = 4
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]


def fnv1a(data):
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def blob_hash(blob):
    return fnv1a(np.ascontiguousarray(blob, dtype="<i4").tobytes())


def analyse(ir):
    """Flatten every process of a local_smart model into writes + rounds of list operations.

    -> dict(nproc, offsets, classes, cls_of, member_of, events=[dict(name, anchor_n, writes, rounds)])
    with rounds = [[op, ...]], op = (kind, q, anchor_off_id, [(off_id, n, mask), ...])."""
    if ir["backend"] != "local_smart":
        raise Unsupported("specialised CUDA is generated for the local_smart backend")
    nproc = len(ir["procs"])
    proc_anchor = tables.proc_anchor_types(ir)
    if any(a == 0 for a in proc_anchor):
        raise Unsupported("a process is registered on several site types")
    if nproc > 64:
        raise Unsupported("more than 64 processes")
    if len(ir["species"]) > 16:
        raise Unsupported("more than 16 species")
    if ir.get("null_species", -1) >= 0:
        raise Unsupported("multi-lattice model (null species on the lattice)")
    classes, cls_of, member_of = devtables.exclusivity_classes(ir, proc_anchor)
    if len(classes) > 255:
        raise Unsupported("more than 255 exclusivity classes")
    offsets = {}

    def off_id(o):
        key = (o[0], o[1], o[2])
        for d in key:
            if not -128 <= d <= 127:
                raise Unsupported("offset out of byte range")
        if key not in offsets:
            if len(offsets) == 127:
                raise Unsupported("more than 127 distinct neighbour offsets")
            offsets[key] = len(offsets)
        return offsets[key]

    off_id([0, 0, 0])
    events = []
    for p in range(nproc):
        base_n, writes, ops = devtables.flatten_event(ir, p)
        if base_n != proc_anchor[p]:
            raise Unsupported("process %d is selected on site type %d but registered on %d"
                              % (p + 1, base_n, proc_anchor[p]))
        if len(writes) > 4:
            raise Unsupported("event writes %d sites" % len(writes))
        for _k, q, aoff, cs, _g in ops:
            if aoff[3] != proc_anchor[q - 1]:
                raise Unsupported("anchor site type mismatch")
            if len(cs) > MAX_COND:
                raise Unsupported("add with %d dynamic probes" % len(cs))
        rounds = devtables.schedule_rounds(ops, lambda op: [op[1]],
                                           lambda op: (cls_of[op[1]], op[2][0], op[2][1], op[2][2]))
        ev_rounds = []
        for rnd in rounds:
            ev_rounds.append([(ops[i][0], ops[i][1], off_id(ops[i][2]),
                               [(off_id(s), s[3], m) for s, m in ops[i][3]]) for i in rnd])
        events.append({"name": ir["procs"][p], "anchor_n": base_n,
                       "writes": [(off_id(o), o[3], old, new) for o, old, new in writes], "rounds": ev_rounds})
    off_list = [None] * len(offsets)
    for key, i in offsets.items():
        off_list[i] = key
    return {"nproc": nproc, "offsets": off_list, "classes": classes, "cls_of": cls_of, "member_of": member_of,
            "events": events, "spuck": ir["spuck"], "n_species": len(ir["species"]),
            "dim": ir["model_dimension"], "species": ir["species"]}


KINDS = {(True, False): "del", (False, True): "add", (True, True): "mixed"}
# estimated SASS instructions of a round body; the event code of the unrolled style must stay inside the SM's
# instruction cache next to the ~700 instructions of the step skeleton (measured: mini_101 and AB profit from
# unrolling, RuO2's 12 k instructions thrash the cache -- DESIGN.md 4.1b)
UNROLL_BUDGET = 1600


def _round_cost(has_del, has_add, nc):
    return (60 if has_del and has_add else 40 if has_add else 35) + 7 * nc


def layout(an):
    """Number the ops and rounds in table order and place the table's regions:
    [A: one uint4 per op][B: bw probe words per op][round words][event rows].  Idle lanes of a round read up
    to 31 entries past its last op, hence the padding behind A and B."""
    rounds = []
    first = 0
    ncmax = 0
    for e, ev in enumerate(an["events"]):
        for r, rnd in enumerate(ev["rounds"]):
            if len(rnd) > 32:
                raise Unsupported("a round with more than 32 ops")
            nc = max(len(op[3]) for op in rnd)
            ncmax = max(ncmax, nc)
            rounds.append({"event": e, "round": r, "count": len(rnd), "nc": nc, "first_op": first,
                           "has_add": any(op[0] == KIND_ADD for op in rnd),
                           "has_del": any(op[0] != KIND_ADD for op in rnd)})
            first += len(rnd)
    n_ops = first
    if n_ops + 32 > 0xffff:
        raise Unsupported("more than 65503 list operations")
    if len(rounds) > 0 and max(len(ev["rounds"]) for ev in an["events"]) > 255:
        raise Unsupported("more than 255 rounds in one event")
    bw = {0: 0, 1: 1, 2: 2, 3: 4, 4: 4}[ncmax]
    off_b = 16 * (n_ops + 32)
    off_rd = (off_b + 4 * bw * (n_ops + 32) + 15) // 16 * 16
    off_ev = (off_rd + 4 * (len(rounds) + 1) + 15) // 16 * 16
    ops_bytes = (off_ev + 24 * an["nproc"] + 15) // 16 * 16
    # round bodies the model uses, the statically most frequent first
    variants = {}
    for rd in rounds:
        key = (rd["has_del"], rd["has_add"], rd["nc"])
        variants[key] = variants.get(key, 0) + 1
    order = sorted(variants, key=lambda k: -variants[k])
    for rd in rounds:
        rd["kind"] = order.index((rd["has_del"], rd["has_add"], rd["nc"]))
    cost = sum(_round_cost(rd["has_del"], rd["has_add"], rd["nc"]) for rd in rounds) + 25 * an["nproc"]
    return {"rounds": rounds, "n_ops": n_ops, "bw": bw, "off_b": off_b, "off_rd": off_rd, "off_ev": off_ev,
            "ops_bytes": ops_bytes, "variants": order, "unrolled_cost": cost}


def generate(ir, blob=None, name=None, max_threads=768, style="auto"):
    """-> (CUDA source text, info dict) for a local_smart model IR.  style: "unrolled" (every process its own
    straight-line case), "compact" (events as descriptor rows over the model's round bodies) or "auto"."""
    if blob is None:
        blob, _info = tables.build_blob(ir)
    an = analyse(ir)
    lay = layout(an)
    rounds = lay["rounds"]
    if style == "auto":
        style = os.environ.get("KMOS_B200_GEN_STYLE") or ("unrolled" if lay["unrolled_cost"] <= UNROLL_BUDGET else "compact")
    if style not in ("unrolled", "compact"):
        raise ValueError("style: unrolled, compact or auto")
    fx = ir.get("fixture")
    name = name or ir.get("model_name") or (fx.get("model") if isinstance(fx, dict) else fx) or "model"
    ident = "".join(ch if ch.isalnum() else "_" for ch in name)
    h = blob_hash(blob)
    P = an["nproc"]
    out = []
    w = out.append
    w("// proclist_%s.cu -- generated by kmos_b200.codegen (version %d, %s style); do not edit." % (
        ident, GEN_VERSION, style))
    w("// CUDA twin of the proclist.f90 `kmos export -b local_smart` writes for this model: run_proc_nr as a switch,")
    w("// every process' replace_species calls and guarded del_proc / if-tree add_proc calls as straight-line code")
    w("// (kmos/io/__init__.py:305-465, 2219-2409, 2568-2655).  model blob hash %016x" % h)
    w("#include \"kb_gen.cuh\"")
    w("")
    w("namespace {")
    w("struct KbModel {")
    w("    static constexpr int P = %d, SPUCK = %d, NOFF = %d, MAX_THREADS = %d;" % (
        P, an["spuck"], len(an["offsets"]), max_threads))
    w("    static constexpr int BW = %d, OFF_B = %d, OFF_EV = %d;  // operand table layout" % (
        lay["bw"], lay["off_b"], lay["off_ev"]))
    w("    typedef KbGenCtx<KbModel> Ctx;")
    by_event = {}
    for rd in rounds:
        by_event.setdefault(rd["event"], []).append(rd)

    def describe(e, ev):
        w("")
        w("    // process %d: %s" % (e + 1, ev["name"]))
        for off, n, old, new in ev["writes"]:
            w("    //   replace_species(site + (%d,%d,%d) type %d, %s -> %s)" % (
                an["offsets"][off] + (n, an["species"][old], an["species"][new])))

    def ops_text(ops):
        return " ".join(("+" if op[0] == KIND_ADD else "-") + str(op[1]) for op in ops)

    if style == "unrolled":
        for e, ev in enumerate(an["events"]):
            describe(e, ev)
            w("    static __device__ __forceinline__ void ev_%d(Ctx& c, const int k) {" % (e + 1))
            w("        c.select<%d>(k);" % e)
            for i, (off, n, old, new) in enumerate(ev["writes"]):
                w("        c.write<%d, %d, %d, %d, %d>();" % (i, off, n, old, new))
            rds = by_event.get(e, [])

            def load(j):
                w("        const uint4 a%d = c.ldA(%d); const KbGenOpB b%d = c.ldB(%d);" % (
                    j, rds[j]["first_op"], j, rds[j]["first_op"]))

            if rds:
                load(0)
            for j, rd in enumerate(rds):
                if j + 1 < len(rds):
                    load(j + 1)
                w("        c.round<%s, %s, %d>(%d, a%d, b%d);  // %s" % (
                    "true" if rd["has_del"] else "false", "true" if rd["has_add"] else "false", rd["nc"],
                    rd["count"], j, j, ops_text(ev["rounds"][j])))
            if not rds:
                w("        __syncwarp();")
            w("    }")
        w("")
        w("    static __device__ __forceinline__ void run_event(Ctx& c, const int pidx, const int k) {")
        w("        switch (pidx) {")
        for e in range(P):
            w("        case %d: ev_%d(c, k); break;" % (e, e + 1))
        w("        default: break;")
        w("        }")
        w("    }")
    else:
        for e, ev in enumerate(an["events"]):
            describe(e, ev)
            for j, rd in enumerate(by_event.get(e, [])):
                w("    //   round %d (%s, %d probes): %s" % (
                    j, KINDS[(rd["has_del"], rd["has_add"])], rd["nc"], ops_text(ev["rounds"][j])))
        w("")
        w("    // the round bodies this model uses; kind = position in this switch")
        w("    static __device__ __forceinline__ void dispatch(Ctx& c, const uint32_t kind, const int count,")
        w("                                                    const uint4 a, const KbGenOpB& b) {")
        if len(lay["variants"]) <= 1:
            for hd, ha, nc in lay["variants"]:
                w("        c.round<%s, %s, %d>(count, a, b);" % ("true" if hd else "false", "true" if ha else "false", nc))
            if not lay["variants"]:
                w("        __syncwarp();")
        else:
            w("        switch (kind) {")
            for i, (hd, ha, nc) in enumerate(lay["variants"]):
                w("        %s c.round<%s, %s, %d>(count, a, b); break;  // %s" % (
                    "default:" if i == len(lay["variants"]) - 1 else "case %d:" % i,
                    "true" if hd else "false", "true" if ha else "false", nc, KINDS[(hd, ha)]))
            w("        }")
        w("    }")
        w("    static __device__ __forceinline__ void run_event(Ctx& c, const int pidx, const int k) {")
        w("        kb_gen_run_compact<KbModel>(c, pidx, k);")
        w("    }")
    w("};")
    w("")
    w("const KbGenOpDesc kb_ops[] = {")
    n_ops = 0
    for e, ev in enumerate(an["events"]):
        for rnd in ev["rounds"]:
            for kind, q, aoff, cs in rnd:
                co = [c[0] for c in cs] + [0] * (MAX_COND - len(cs))
                cn = [c[1] for c in cs] + [0] * (MAX_COND - len(cs))
                cm = [c[2] for c in cs] + [0] * (MAX_COND - len(cs))
                w("    {%d, %d, %d, %d, %d, %d, {%s}, {%s}, {%s}}," % (
                    1 if kind == KIND_ADD else 0, q - 1, an["cls_of"][q], an["member_of"][q], aoff, len(cs),
                    ",".join(map(str, co)), ",".join(map(str, cn)), ",".join(map(str, cm))))
                n_ops += 1
    assert n_ops == lay["n_ops"]
    if n_ops == 0:
        w("    {0, 0, 0, 0, 0, 0, {0,0,0,0}, {0,0,0,0}, {0,0,0,0}},")
    w("};")
    w("const KbGenRoundDesc kb_rounds[] = {")
    for rd in rounds:
        w("    {%d, %d, %d, %d}," % (rd["first_op"], rd["count"], rd["nc"], rd["kind"]))
    if not rounds:
        w("    {0, 0, 0, 0},")
    w("};")
    w("const KbGenEventDesc kb_events[] = {")
    fr = 0
    for ev in an["events"]:
        w("    {%d, %d}," % (fr, len(ev["rounds"])))
        fr += len(ev["rounds"])
    w("};")
    w("const int8_t kb_offsets[] = {%s};" % ", ".join("%d,%d,%d" % o for o in an["offsets"]))
    w("const uint32_t kb_writes[] = {")
    for ev in an["events"]:
        ws = [(off | (n << 8) | (old << 16) | (new << 24)) for off, n, old, new in ev["writes"]]
        ws += [0] * (4 - len(ws))
        w("    %s," % ", ".join("0x%08xu" % x for x in ws))
    w("};")
    w("const uint8_t kb_proc_cls[] = {%s};" % ", ".join(str(an["cls_of"][q]) for q in range(1, P + 1)))
    w("const uint8_t kb_proc_member[] = {%s};" % ", ".join(str(an["member_of"][q]) for q in range(1, P + 1)))
    w("const KbGenInfo kb_info = {KB_GEN_ABI, %d, %d, %d, %d, %d, %d, %d, %d," % (
        P, an["n_species"], an["spuck"], an["dim"], len(an["offsets"]), len(an["classes"]), n_ops, len(rounds)))
    w("                           %d, %d, %d, %d, %d, %d, %d, 0x%016xull, \"%s\"," % (
        lay["bw"], lay["off_b"], lay["off_rd"], lay["off_ev"], lay["ops_bytes"], max_threads,
        1 if style == "compact" else 0, h, ident))
    w("                           kb_ops, kb_rounds, kb_events, kb_offsets, kb_writes, kb_proc_cls, kb_proc_member};")
    w("}  // namespace")
    w("")
    w("KB_GEN_MODULE(KbModel, kb_info)")
    w("")
    info = {"name": ident, "hash": h, "n_ops": n_ops, "n_rounds": len(rounds), "ops_bytes": lay["ops_bytes"],
            "n_classes": len(an["classes"]), "n_offsets": len(an["offsets"]), "style": style,
            "unrolled_cost": lay["unrolled_cost"], "variants": lay["variants"],
            "rounds_per_event": [len(ev["rounds"]) for ev in an["events"]]}
    return "\n".join(out), info


def write_source(ir, out_dir, blob=None, name=None, max_threads=768, style="auto"):
    """Write proclist_<model>.cu into `out_dir` (next to the exported Fortran); returns its path."""
    src, info = generate(ir, blob, name, max_threads, style)
    path = os.path.join(out_dir, "proclist_%s.cu" % info["name"])
    with open(path, "w") as f:
        f.write(src)
    return path, info


def _skeleton_digest():
    hsh = hashlib.sha256()
    for f in ("kb_gen.cuh", "kb_smem.cuh", "kb_common.h"):
        with open(os.path.join(CSRC, f), "rb") as fh:
            hsh.update(fh.read())
    return hsh


def build(ir, blob=None, name=None, out_dir=None, max_threads=768, verbose=False, style="auto"):
    """Generate and compile; the shared object is cached under the content hash of source + skeleton.

    -> path of proclist_<model>_<hash>.so"""
    src, info = generate(ir, blob, name, max_threads, style)
    hsh = _skeleton_digest()
    hsh.update(src.encode())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    tag = hsh.hexdigest()[:16]
    out_dir = out_dir or CACHE
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "proclist_%s_%s.so" % (info["name"], tag))
    if os.path.exists(so):
        return so
    cu = os.path.join(out_dir, "proclist_%s_%s.cu" % (info["name"], tag))
    with open(cu, "w") as f:
        f.write(src)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", CSRC, "-o", so + ".tmp", cu]
    subprocess.check_call(cmd)
    os.replace(so + ".tmp", so)
    return so


def find_built(ir, blob=None, name=None, out_dir=None, max_threads=768, style="auto"):
    """Path of the cached shared object for this model, or None (no compiler is invoked)."""
    src, info = generate(ir, blob, name, max_threads, style)
    hsh = _skeleton_digest()
    hsh.update(src.encode())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    so = os.path.join(out_dir or CACHE, "proclist_%s_%s.so" % (info["name"], hsh.hexdigest()[:16]))
    return so if os.path.exists(so) else None


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    do_build = "--build" in sys.argv
    style = "auto"
    if "--style" in sys.argv:
        style = sys.argv[sys.argv.index("--style") + 1]
        args = [a for a in args if a != style]
    out = None
    if "-o" in sys.argv:
        out = sys.argv[sys.argv.index("-o") + 1]
        args = [a for a in args if a != out]
    src_path = args[0]
    if os.path.isdir(src_path):
        src_path = os.path.join(src_path, "model_tables.json")
    model_ir = tables.load_ir(src_path)
    out = out or os.path.dirname(os.path.abspath(src_path))
    path, inf = write_source(model_ir, out, style=style)
    print("%s: %s style, %d ops in %d rounds, %d table bytes" % (path, inf["style"], inf["n_ops"], inf["n_rounds"],
                                                             inf["ops_bytes"]))
    if do_build:
        print(build(model_ir, verbose=True, style=style))
