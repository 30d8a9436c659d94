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

    python -m kmos_b200.codegen <model_tables.json | export_dir> [-o out_dir] [--build]
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
GEN_VERSION = 2
MAX_COND = 4
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


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


def _qw(nc):
    return 0 if nc == 0 else (1 + 3 * nc + 3) // 4


def layout(an):
    """Place every round's operand arrays: region A (one uint4 per op, all rounds), then one region per probe
    count with equally sized entries -- idle lanes of a round read on into entries of the same format."""
    a_off = 0
    rounds = []
    for e, ev in enumerate(an["events"]):
        for r, rnd in enumerate(ev["rounds"]):
            nc = max(len(op[3]) for op in rnd)
            rounds.append({"event": e, "round": r, "count": len(rnd), "nc": nc, "a_off": a_off,
                           "has_add": any(op[0] == KIND_ADD for op in rnd),
                           "has_del": any(op[0] != KIND_ADD for op in rnd)})
            a_off += 16 * len(rnd)
    pos = a_off + 512
    for qw in range(0, _qw(MAX_COND) + 1):
        stride = 4 if qw == 0 else 16 * qw
        used = False
        for rd in rounds:
            if _qw(rd["nc"]) == qw:
                rd["b_off"] = pos
                pos += stride * rd["count"]
                used = True
        if used:
            pos = (pos + 32 * stride + 15) // 16 * 16
    first = 0
    for rd in rounds:
        rd["first_op"] = first
        first += rd["count"]
    return rounds, pos


def generate(ir, blob=None, name=None, max_threads=768):
    """-> (CUDA source text, info dict) for a local_smart model IR."""
    if blob is None:
        blob, _info = tables.build_blob(ir)
    an = analyse(ir)
    rounds, ops_bytes = layout(an)
    fx = ir.get("fixture")
    name = name or ir.get("model_name") or (fx.get("model") if isinstance(fx, dict) else fx) or "model"
    ident = "".join(ch if ch.isalnum() else "_" for ch in name)
    h = blob_hash(blob)
    P = an["nproc"]
    out = []
    w = out.append
    w("// proclist_%s.cu -- generated by kmos_b200.codegen (version %d); do not edit." % (ident, GEN_VERSION))
    w("// CUDA twin of the proclist.f90 `kmos export -b local_smart` writes for this model: run_proc_nr as a switch,")
    w("// every process' replace_species calls and guarded del_proc / if-tree add_proc calls as straight-line code")
    w("// (kmos/io/__init__.py:305-465, 2219-2409, 2568-2655).  model blob hash %016x" % h)
    w("#include \"kb_gen.cuh\"")
    w("")
    w("namespace {")
    w("struct KbModel {")
    w("    static constexpr int P = %d, SPUCK = %d, NOFF = %d, MAX_THREADS = %d;" % (
        P, an["spuck"], len(an["offsets"]), max_threads))
    w("    typedef KbGenCtx<KbModel> Ctx;")
    by_event = {}
    for rd in rounds:
        by_event.setdefault(rd["event"], []).append(rd)
    for e, ev in enumerate(an["events"]):
        w("")
        w("    // process %d: %s" % (e + 1, ev["name"]))
        for i, (off, n, old, new) in enumerate(ev["writes"]):
            w("    //   replace_species(site + (%d,%d,%d) type %d, %s -> %s)" % (
                an["offsets"][off] + (n, an["species"][old], an["species"][new])))
        w("    static __device__ __forceinline__ void ev_%d(Ctx& c, const int k) {" % (e + 1))
        w("        c.select<%d>(k);" % e)
        for i, (off, n, old, new) in enumerate(ev["writes"]):
            w("        c.write<%d, %d, %d, %d, %d>();" % (i, off, n, old, new))
        rds = by_event.get(e, [])

        def load(j):
            rd = rds[j]
            w("        const uint4 a%d = c.ldA<%d>(); const KbGenOpB b%d = c.ldB<%d, %d>();" % (
                j, rd["a_off"], j, rd["b_off"], rd["nc"]))

        if rds:
            load(0)
        for j, rd in enumerate(rds):
            if j + 1 < len(rds):
                load(j + 1)
            ops = ev["rounds"][j]
            w("        c.round<%d, %s, %s, %d>(a%d, b%d);  // %s" % (
                rd["count"], "true" if rd["has_del"] else "false", "true" if rd["has_add"] else "false", rd["nc"],
                j, j, " ".join(("+" if op[0] == KIND_ADD else "-") + str(op[1]) for op in ops)))
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
    if n_ops == 0:
        w("    {0, 0, 0, 0, 0, 0, {0,0,0,0}, {0,0,0,0}, {0,0,0,0}},")
    w("};")
    w("const KbGenRoundDesc kb_rounds[] = {")
    for rd in rounds:
        w("    {%d, %d, %d, %d, %d}," % (rd["first_op"], rd["count"], rd["nc"], rd["a_off"], rd["b_off"]))
    if not rounds:
        w("    {0, 0, 0, 0, 0},")
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
    w("const KbGenInfo kb_info = {KB_GEN_ABI, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, 0x%016xull, \"%s\"," % (
        P, an["n_species"], an["spuck"], an["dim"], len(an["offsets"]), len(an["classes"]), n_ops, len(rounds),
        ops_bytes, max_threads, h, ident))
    w("                           kb_ops, kb_rounds, kb_offsets, kb_writes, kb_proc_cls, kb_proc_member};")
    w("}  // namespace")
    w("")
    w("KB_GEN_MODULE(KbModel, kb_info)")
    w("")
    info = {"name": ident, "hash": h, "n_ops": n_ops, "n_rounds": len(rounds), "ops_bytes": ops_bytes,
            "n_classes": len(an["classes"]), "n_offsets": len(an["offsets"]),
            "rounds_per_event": [len(ev["rounds"]) for ev in an["events"]]}
    return "\n".join(out), info


def write_source(ir, out_dir, blob=None, name=None, max_threads=768):
    """Write proclist_<model>.cu into `out_dir` (next to the exported Fortran); returns its path."""
    src, info = generate(ir, blob, name, max_threads)
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


def build(ir, blob=None, name=None, out_dir=None, max_threads=768, verbose=False):
    """Generate and compile; the shared object is cached under the content hash of source + skeleton.

    -> path of proclist_<model>_<hash>.so"""
    src, info = generate(ir, blob, name, max_threads)
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


def find_built(ir, blob=None, name=None, out_dir=None, max_threads=768):
    """Path of the cached shared object for this model, or None (no compiler is invoked)."""
    src, info = generate(ir, blob, name, max_threads)
    hsh = _skeleton_digest()
    hsh.update(src.encode())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    so = os.path.join(out_dir or CACHE, "proclist_%s_%s.so" % (info["name"], hsh.hexdigest()[:16]))
    return so if os.path.exists(so) else None


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    do_build = "--build" in sys.argv
    out = None
    if "-o" in sys.argv:
        out = sys.argv[sys.argv.index("-o") + 1]
        args = [a for a in args if a != out]
    src_path = args[0]
    if os.path.isdir(src_path):
        src_path = os.path.join(src_path, "model_tables.json")
    model_ir = tables.load_ir(src_path)
    out = out or os.path.dirname(os.path.abspath(src_path))
    path, inf = write_source(model_ir, out)
    print("%s: %d ops in %d rounds, %d table bytes" % (path, inf["n_ops"], inf["n_rounds"], inf["ops_bytes"]))
    if do_build:
        print(build(model_ir, verbose=True))
