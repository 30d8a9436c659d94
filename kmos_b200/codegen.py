"""Exporter back end: emit a model's proclist as specialised CUDA (``proclist_<model>.cu``).

kmos writes ``run_proc_nr`` and the ``put_``/``take_`` routines as model-specific Fortran
(kmos/io/__init__.py:305-465 write_proclist_run_proc_nr_smart, :2219-2409 write_proclist_put_take,
:2568-2655 _write_optimal_iftree); ``export_source`` (:3884-3974, hook point :3958-3973) is where the files
are written.  This module is the CUDA twin of that step for the local_smart backend: from the same rule IR the
Fortran was parsed into (kmos_b200.fortran_ir -- so statement order, and with it the order of ``avail_sites``,
is the Fortran's) it emits

  * the model's compile-time constants (process count, sites per cell, neighbour offsets, probes per add,
    table layout) and the *lane-group width*: how many lanes step one replica (32, 16 or 8).  The rounds of
    an event are bounded by the dependency chains of its list operations, not by the number of lanes, so the
    generator picks the narrowest group that does not lengthen them: the warp then steps 2 or 4 replicas
    with every instruction;
  * per process: its ``replace_species`` calls and its guarded ``del_proc`` / if-tree ``add_proc`` calls
    scheduled into rounds of at most ``lpr`` operations, as static descriptor tables the host turns into the
    geometry-specialised operand table (``kb_gen_fill_tables`` in csrc/kb_gen.cuh) -- events are rows of a
    table and not unrolled cases because the replicas of a warp (and the warps of an SM) are in different
    events at any time: measured, RuO2's 36 unrolled cases (196 KB of SASS) thrash the instruction cache.

The result is compiled with nvcc for sm_100a into ``proclist_<model>_<hash>.so`` (cached by content hash) and
attached to a batch with ``kmos_b200_batch_attach_proclist``; the table interpreter (kb_smem.cuh) stays the
path for models this generator declines (``Unsupported``).

    python -m kmos_b200.codegen <model_tables.json | export_dir> [-o out_dir] [--lpr 8|16|32] [--build]
"""
import hashlib
import os
import subprocess
import sys
import threading

import numpy as np

from . import devtables, tables
from .devtables import KIND_ADD, Unsupported

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CACHE = os.path.join(HERE, "_proclist_cache")
GEN_VERSION = 4
MAX_COND = 4
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]
# threads per CTA the kernel is compiled for (register budget 65536 / threads): narrow groups run fewer warps
MAX_THREADS = {32: 768, 16: 640, 8: 512, 4: 512}


def fnv1a(data):
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def blob_hash(blob):
    return fnv1a(np.ascontiguousarray(blob, dtype="<i4").tobytes())


KINDS = {(True, False): "del", (False, True): "add", (True, True): "mixed"}


def _flatten(ir):
    """Every process of a local_smart model as writes + list operations in the Fortran's textual order.

    -> dict(nproc, offsets, classes, cls_of, member_of, events=[dict(name, writes, ops)]) with
    op = (kind, q, anchor_off_id, [(off_id, n, mask), ...], group)."""
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
    if ir["spuck"] > 128:
        raise Unsupported("more than 128 sites per cell")
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
        flat = []
        for k, q, aoff, cs, grp in ops:
            if aoff[3] != proc_anchor[q - 1]:
                raise Unsupported("anchor site type mismatch")
            if len(cs) > MAX_COND:
                raise Unsupported("add with %d dynamic probes" % len(cs))
            flat.append((k, q, off_id(aoff), [(off_id(s_), s_[3], m) for s_, m in cs], grp, tuple(aoff[:3])))
        events.append({"name": ir["procs"][p], "anchor_n": base_n,
                       "writes": [(off_id(o), o[3], old, new) for o, old, new in writes], "ops": flat})
    off_list = [None] * len(offsets)
    for key, i in offsets.items():
        off_list[i] = key
    return {"nproc": nproc, "offsets": off_list, "classes": classes, "cls_of": cls_of, "member_of": member_of,
            "events": events, "spuck": ir["spuck"], "n_species": len(ir["species"]),
            "dim": ir["model_dimension"], "species": ir["species"]}


def _schedule(an, lpr):
    """Rounds of every event for groups of `lpr` lanes -> [[[op, ...], ...], ...] (devtables.schedule_rounds)."""
    cls_of = an["cls_of"]
    out = []
    for ev in an["events"]:
        ops = ev["ops"]
        rounds = devtables.schedule_rounds(ops, lambda op: [op[1]], lambda op: (cls_of[op[1]],) + op[5], width=lpr)
        out.append([[ops[i] for i in rnd] for rnd in rounds])
    return out


def choose_lpr(an):
    """Lanes per replica: the narrowest group that lengthens the events' rounds (mean over processes) by less
    than 15 % over a full warp -- the rounds are then set by the dependency chains, and the warp's other lanes
    step further replicas."""
    env = os.environ.get("KMOS_B200_GEN_LPR")
    if env:
        if int(env) not in (4, 8, 16, 32):
            raise ValueError("KMOS_B200_GEN_LPR: 4, 8, 16 or 32")
        return int(env)
    mean = {}
    for lpr in (32, 16, 8):
        rs = _schedule(an, lpr)
        mean[lpr] = sum(len(r) for r in rs) / float(max(1, len(rs)))
    for lpr in (8, 16):
        if mean[lpr] <= 1.15 * mean[32] + 1e-9:
            return lpr
    return 32


def expected_max_rounds(rounds_per_event, n_groups):
    """Expected number of rounds of the slowest of `n_groups` events drawn uniformly from the model's processes:
    the groups of a warp run their rounds in lock step."""
    r = np.asarray(rounds_per_event, dtype=float)
    if r.size == 0:
        return 0.0
    vals = np.unique(r)
    cdf = np.array([(r <= v).mean() for v in vals]) ** n_groups
    return float((vals * np.diff(np.concatenate([[0.0], cdf]))).sum())


def lane_group_widths(n_proc):
    """Candidate lanes-per-replica of a model: groups of four lanes (eight replicas per warp) only where a lane
    then owns at most four processes (registers) -- they pay for models with one or two ops per event
    (mini_101: +23 % over eight lanes)."""
    return [4, 8, 16, 32] if n_proc <= 16 else [8, 16, 32]


def lane_group_score(n_proc, emax, n_groups, replicas_per_sm):
    """Relative throughput estimate of a lane-group width for one batch geometry (higher is better).

    A group-step costs about I = 170 + 3.5 P + 100 E[max rounds] warp instructions.  With `replicas_per_sm`
    resident replicas in w = replicas / groups warps the kernel is bound either by the latency of that chain --
    2.3 + 0.2 w cycles per instruction (issue contention grows with the resident warps; below 8 warps per SM a
    scheduler has nothing to overlap, hence the w/8 factor), every resident replica advancing one step per
    chain -- or by issue slots (about 3 per cycle, a warp instruction serving 32/lpr replicas).  Fitted on
    RuO2 (full and 2048-replica batches), ZGB, AB, mini_101 and pairwise at 8/16/32 lanes: it
    picks the measured best width in all six cases."""
    instr = 170.0 + 3.5 * n_proc + 100.0 * emax
    warps = replicas_per_sm / float(n_groups)
    latency = replicas_per_sm / (2.3 + 0.2 * warps) * min(1.0, warps / 8.0)
    return min(latency, 3.0 * n_groups) / instr


def analyse(ir, lpr=None):
    """_flatten + the lane-group width + every event's rounds (ev["rounds"] = [[op, ...], ...])."""
    an = _flatten(ir)
    if lpr and int(lpr) not in lane_group_widths(an["nproc"]):
        raise Unsupported("groups of %d lanes per replica with %d processes" % (int(lpr), an["nproc"]))
    an["lpr"] = lpr or choose_lpr(an)
    for ev, rounds in zip(an["events"], _schedule(an, an["lpr"])):
        ev["rounds"] = rounds
    return an


def layout(an):
    """Number the ops and rounds in table order and place the table's regions:
    [A: one word per op][B: bw probe words per op][round words][event rows][write rows].  Idle lanes of a
    round read up to lpr-1 entries past its last op, hence the padding behind A and B."""
    rounds = []
    first = 0
    ncmax = 0
    lpr = an["lpr"]
    for e, ev in enumerate(an["events"]):
        for r, rnd in enumerate(ev["rounds"]):
            if len(rnd) > lpr:
                raise Unsupported("a round with more than %d ops" % lpr)
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
    off_b = (4 * (n_ops + 32) + 15) // 16 * 16
    off_rd = (off_b + 4 * bw * (n_ops + 32) + 15) // 16 * 16
    off_ev = (off_rd + 4 * (len(rounds) + 1) + 15) // 16 * 16
    off_wr = off_ev + 16 * an["nproc"]
    ops_bytes = off_wr + 16 * an["nproc"]
    return {"rounds": rounds, "n_ops": n_ops, "bw": bw, "off_b": off_b, "off_rd": off_rd, "off_ev": off_ev,
            "off_wr": off_wr, "ops_bytes": ops_bytes}


def generate(ir, blob=None, name=None, lpr=None):
    """-> (CUDA source text, info dict) for a local_smart model IR.  lpr: lanes per replica (default: chosen
    from the model's rounds, or KMOS_B200_GEN_LPR)."""
    if blob is None:
        blob, _info = tables.build_blob(ir)
    an = analyse(ir, lpr)
    lay = layout(an)
    rounds = lay["rounds"]
    lpr = an["lpr"]
    max_threads = int(os.environ.get("KMOS_B200_GEN_MAX_THREADS", MAX_THREADS[lpr]))  # experiments: register budget
    fx = ir.get("fixture")
    name = name or ir.get("model_name") or (fx.get("model") if isinstance(fx, dict) else fx) or "model"
    ident = "".join(ch if ch.isalnum() else "_" for ch in name)
    h = blob_hash(blob)
    P = an["nproc"]
    out = []
    w = out.append
    w("// proclist_%s.cu -- generated by kmos_b200.codegen (version %d); do not edit." % (ident, GEN_VERSION))
    w("// CUDA twin of the proclist.f90 `kmos export -b local_smart` writes for this model: run_proc_nr, every")
    w("// process' replace_species calls and its guarded del_proc / if-tree add_proc calls (kmos/io/__init__.py:")
    w("// 305-465, 2219-2409, 2568-2655), scheduled into rounds for groups of %d lanes per replica." % lpr)
    w("// model blob hash %016x" % h)
    w("#include \"kb_gen.cuh\"")
    w("")
    w("namespace {")
    w("struct KbModel {")
    w("    static constexpr int P = %d, SPUCK = %d, NOFF = %d;" % (P, an["spuck"], len(an["offsets"])))
    w("    static constexpr int LPR = %d, MAX_THREADS = %d;  // lanes per replica, threads per CTA" % (lpr, max_threads))
    w("    static constexpr int BW = %d, OFF_B = %d, OFF_EV = %d, OFF_WR = %d;  // operand table layout" % (
        lay["bw"], lay["off_b"], lay["off_ev"], lay["off_wr"]))
    w("};")
    by_event = {}
    for rd in rounds:
        by_event.setdefault(rd["event"], []).append(rd)

    def ops_text(ops):
        return " ".join(("+" if op[0] == KIND_ADD else "-") + str(op[1]) for op in ops)

    for e, ev in enumerate(an["events"]):
        w("")
        w("// process %d: %s" % (e + 1, ev["name"]))
        for off, n, old, new in ev["writes"]:
            w("//   replace_species(site + (%d,%d,%d) type %d, %s -> %s)" % (
                an["offsets"][off] + (n, an["species"][old], an["species"][new])))
        for j, rd in enumerate(by_event.get(e, [])):
            w("//   round %d (%s, %d probes): %s" % (
                j, KINDS[(rd["has_del"], rd["has_add"])], rd["nc"], ops_text(ev["rounds"][j])))
    w("")
    w("const KbGenOpDesc kb_ops[] = {")
    n_ops = 0
    for e, ev in enumerate(an["events"]):
        for rnd in ev["rounds"]:
            for op in rnd:
                kind, q, aoff, cs = op[0], op[1], op[2], op[3]
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
        w("    {%d, %d, %d, 0}," % (rd["first_op"], rd["count"], rd["nc"]))
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
    w("                           %d, %d, %d, %d, %d, %d, %d, %d, 0x%016xull, \"%s\"," % (
        lay["bw"], lay["off_b"], lay["off_rd"], lay["off_ev"], lay["off_wr"], lay["ops_bytes"], max_threads,
        lpr, h, ident))
    w("                           kb_ops, kb_rounds, kb_events, kb_offsets, kb_writes, kb_proc_cls, kb_proc_member};")
    w("}  // namespace")
    w("")
    w("KB_GEN_MODULE(KbModel, kb_info)")
    w("")
    info = {"name": ident, "hash": h, "n_ops": n_ops, "n_rounds": len(rounds), "ops_bytes": lay["ops_bytes"],
            "n_classes": len(an["classes"]), "n_offsets": len(an["offsets"]), "lpr": lpr,
            "rounds_per_event": [len(ev["rounds"]) for ev in an["events"]]}
    return "\n".join(out), info


def write_source(ir, out_dir, blob=None, name=None, lpr=None):
    """Write proclist_<model>.cu into `out_dir` (next to the exported Fortran); returns its path."""
    src, info = generate(ir, blob, name, lpr)
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


def _extra_defs():
    """-D switches for kernel experiments (KMOS_B200_GEN_DEFS="KB_GEN_X KB_GEN_Y"); part of the cache key."""
    return ["-D" + d for d in os.environ.get("KMOS_B200_GEN_DEFS", "").split()]


def _so_path(src, info, out_dir):
    hsh = _skeleton_digest()
    hsh.update(src.encode())
    hsh.update(" ".join(NVCC_FLAGS + _extra_defs()).encode())
    return os.path.join(out_dir or CACHE, "proclist_%s_%s.so" % (info["name"], hsh.hexdigest()[:16]))


def build(ir, blob=None, name=None, out_dir=None, verbose=False, lpr=None):
    """Generate and compile; the shared object is cached under the content hash of source + skeleton.

    -> path of proclist_<model>_<hash>.so"""
    src, info = generate(ir, blob, name, lpr)
    so = _so_path(src, info, out_dir)
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if os.path.exists(so):
        return so
    cu = so[:-3] + ".cu"
    with open(cu, "w") as f:
        f.write(src)
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = "%s.%d.%d.tmp" % (so, os.getpid(), threading.get_ident())
    cmd = [nvcc] + NVCC_FLAGS + _extra_defs() + (["-Xptxas", "-v"] if verbose else []) + ["-I", CSRC, "-o", tmp, cu]
    subprocess.check_call(cmd)
    os.replace(tmp, so)
    return so


def find_built(ir, blob=None, name=None, out_dir=None, lpr=None):
    """Path of the cached shared object for this model, or None (no compiler is invoked)."""
    src, info = generate(ir, blob, name, lpr)
    so = _so_path(src, info, out_dir)
    return so if os.path.exists(so) else None


if __name__ == "__main__":
    argv = sys.argv[1:]
    do_build = "--build" in argv
    lpr_arg = None
    out = None
    args = []
    i = 0
    while i < len(argv):
        if argv[i] == "--lpr":
            lpr_arg = int(argv[i + 1]); i += 2
        elif argv[i] == "-o":
            out = argv[i + 1]; i += 2
        elif argv[i].startswith("-"):
            i += 1
        else:
            args.append(argv[i]); i += 1
    src_path = args[0]
    if os.path.isdir(src_path):
        src_path = os.path.join(src_path, "model_tables.json")
    model_ir = tables.load_ir(src_path)
    out = out or os.path.dirname(os.path.abspath(src_path))
    os.makedirs(out, exist_ok=True)
    path, inf = write_source(model_ir, out, lpr=lpr_arg)
    print("%s: %d lanes per replica, %d ops in %d rounds, %d table bytes" % (
        path, inf["lpr"], inf["n_ops"], inf["n_rounds"], inf["ops_bytes"]))
    if do_build:
        print(build(model_ir, verbose=True, lpr=lpr_arg))
