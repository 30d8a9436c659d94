"""Per-event lane tables for the shared-memory CUDA kernel (SEC_DEVICE of the model blob).

The generated Fortran executes, for a chosen (process, site), a fixed *sequence* of put_/take_ routines
(local_smart; kmos/io/__init__.py:305-465, 2219-2409).  Each routine is: one ``replace_species``, a flat
list of guarded ``del_proc`` and an if-tree of ``get_species`` probes ending in ``add_proc`` leaves.
Bit-exact parity only requires that, *per process*, the add/del calls hit ``avail_sites`` in the
reference order (different processes own disjoint rows of ``avail_sites``/``nr_of_sites``;
base.mpy:211-302).  So the whole event is flattened here, at export time, into

    writes   the replace_species calls (site offset, old, new), in order
    ops      every del/add candidate of every action, in textual (= execution) order:
             DEL_IF (q, anchor)            -- guard avail_sites(q, anchor, 2) /= 0 read at execution time
             ADD    (q, anchor, conds...)  -- leaf of the if-tree; conds = species tests on the path

and then scheduled into *rounds*: within a round every op touches a different process, so the 32 lanes
of the replica's warp execute a round concurrently; ops of the same process land in successive rounds in
their original order.

Static resolution: while an event runs, the species on its own action sites are known at compile time
(the event's conditions fix them before, its actions after), so every if-tree probe of such a site is
folded here -- ops that cannot fire are dropped, the rest only probe sites the event does not modify.
Lattice probes are therefore independent of the order of the event's own writes and all lanes can
evaluate them up front.

Layout of the int32 section (all offsets relative to the section start):
    [0] version  [1] supported  [2] n_events  [3] events_off  [4] ops_off  [5] n_ops
    [6] anchors_off [7] n_anchors  [8] conds_off  [9] n_conds  [10] max_rounds  [11] max_ops_per_event
    [12] proc_anchor_off (n_proc words: 1-based site type each process is registered on)
    events: EVENT_STRIDE words each:
        [0] ops_start [1] n_rounds [2] n_writes [3] base site type
        [4..4+MAX_ROUNDS)  cumulative op count at the end of each round
        then MAX_WRITES x (packed offset, old | new<<8)
    ops: 2 words:  w0 = kind | q<<4 | anchor_idx<<16 | ncond<<24 ;  w1 = 4 x u8 cond indices
    anchors / cond sites: 1 word: (dx&255) | (dy&255)<<8 | (dz&255)<<16 | n<<24   (n = absolute site type)
    conds: 2 words: packed site, species mask
"""
MAX_ROUNDS = 8
MAX_WRITES = 4
EVENT_STRIDE = 4 + MAX_ROUNDS + 2 * MAX_WRITES
KIND_NOP, KIND_DEL_IF, KIND_ADD = 0, 1, 2
DEV_VERSION = 1
WARP = 32


class Unsupported(Exception):
    pass


def pack_site(off):
    dx, dy, dz, n = off
    for d in (dx, dy, dz):
        if not -128 <= d <= 127:
            raise Unsupported("offset out of byte range")
    if not 0 < n < 128:
        raise Unsupported("site type out of range")
    return (dx & 255) | ((dy & 255) << 8) | ((dz & 255) << 16) | (n << 24)


def _add4(a, b):
    return [a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]]


def flatten_event(ir, proc_index):
    """-> (base_n, writes, ops) for process `proc_index` (0-based) of a local_smart model.

    writes: [(off4_abs, old, new)]; ops: [(kind, q, anchor_off4_abs, [(off4_abs, mask)...])].
    Offsets are relative to the selected site's cell, 4th component = absolute site type.
    """
    calls = ir["run_proc"][proc_index]
    rsite = ir["routine_site"]
    all_mask = (1 << len(ir["species"])) - 1
    # site type of the event's base coordinate: routine's site type minus the call's dn
    first = calls[0]
    base_n = rsite[first[1]] - first[2][3]
    # species known on the event's own sites: before the first write = that write's `old`
    known = {}
    seq = []  # (routine stmts, abs offset of routine base)
    for _c, rname, off in calls:
        rb = [off[0], off[1], off[2], base_n + off[3]]
        if rb[3] != rsite[rname]:
            raise Unsupported("inconsistent site type in run_proc of process %d" % (proc_index + 1))
        seq.append((ir["routines"][rname], rb))
    for stmts, rb in seq:
        for st in stmts:
            if st[0] == "replace":
                key = tuple(_add4(rb, st[1]))
                known.setdefault(key, st[2])
    writes, ops = [], []

    def walk(block, rb, conds):
        for st in block:
            k = st[0]
            if k == "replace":
                key = tuple(_add4(rb, st[1]))
                if known.get(key) != st[2]:
                    raise Unsupported("replace_species old-species mismatch at compile time")
                known[key] = st[3]
                writes.append((list(key), st[2], st[3]))
            elif k == "if_can":
                body = st[3]
                if len(body) != 1 or body[0][0] != "del" or body[0][1] != st[1] or body[0][2] != st[2]:
                    raise Unsupported("if_can body is not the matching del_proc")
                if conds:
                    raise Unsupported("guarded del inside a select")
                ops.append((KIND_DEL_IF, st[1], _add4(rb, st[2]), []))
            elif k == "add":
                if isinstance(st[1], list) or st[3] is not None:
                    raise Unsupported("nli/otf add in local_smart table")
                ops.append((KIND_ADD, st[1], _add4(rb, st[2]), list(conds)))
            elif k == "select":
                site = _add4(rb, st[1])
                seen_mask = 0
                for key, body in st[2]:
                    if key is None:
                        mask = all_mask & ~seen_mask
                    else:
                        mask = 0
                        for s in key:
                            mask |= 1 << s
                        mask &= ~seen_mask  # select case executes the first match only
                    seen_mask |= mask
                    if tuple(site) in known:
                        if (mask >> known[tuple(site)]) & 1:
                            walk(body, rb, conds)
                    else:
                        walk(body, rb, conds + [(site, mask)])
            elif k == "del":
                raise Unsupported("unguarded del in local_smart table")
            else:
                raise Unsupported("statement %r in put/take routine" % k)

    for stmts, rb in seq:
        walk(stmts, rb, [])
    return base_n, writes, ops


def schedule_rounds(ops, resource_of=None):
    """Greedy list scheduling: ops of one process keep their order in successive rounds, a round holds at
    most WARP ops.  Returns list of rounds (lists of op indices)."""
    rounds = []
    last_round = {}
    for i, op in enumerate(ops):
        res = op[1] if resource_of is None else resource_of[op[1]]
        r = last_round.get(res, -1) + 1
        while True:
            if r == len(rounds):
                rounds.append([])
            if len(rounds[r]) < WARP:
                break
            r += 1
        rounds[r].append(i)
        last_round[res] = r
    return rounds


def compile_device_tables(ir, asm=None):
    info = {"supported": False}
    nproc = len(ir["procs"])
    header = [DEV_VERSION, 0] + [0] * 11
    if ir["backend"] != "local_smart":
        info["reason"] = "shared-memory tables are generated for local_smart only"
        return header, info
    try:
        from .tables import proc_site_masks
        masks = proc_site_masks(ir)
        proc_anchor = []
        for m in masks:
            if m == 0 or (m & (m - 1)):
                raise Unsupported("process registered on %s site types" % bin(m).count("1"))
            proc_anchor.append(m.bit_length())
        if nproc > 64:
            raise Unsupported("more than 64 processes")
        if len(ir["species"]) > 16:
            raise Unsupported("more than 16 species")
        anchors, conds = {}, {}
        ops_words, events_words = [], []
        stats = []
        max_rounds = max_ops = 0
        for p in range(nproc):
            base_n, writes, ops = flatten_event(ir, p)
            if base_n != proc_anchor[p]:
                raise Unsupported("process %d is selected on site type %d but registered on %d"
                                  % (p + 1, base_n, proc_anchor[p]))
            for _k, q, aoff, _c in ops:
                if aoff[3] != proc_anchor[q - 1]:
                    raise Unsupported("anchor site type mismatch")
            rounds = schedule_rounds(ops)
            if len(rounds) > MAX_ROUNDS:
                raise Unsupported("event needs %d rounds" % len(rounds))
            if len(writes) > MAX_WRITES:
                raise Unsupported("event writes %d sites" % len(writes))
            ops_start = len(ops_words) // 2
            cum = []
            n = 0
            for rnd in rounds:
                for i in rnd:
                    kind, q, aoff, cs = ops[i]
                    if len(cs) > 4:
                        raise Unsupported("add with %d dynamic conditions" % len(cs))
                    a_idx = anchors.setdefault(pack_site(aoff), len(anchors))
                    c_idx = [conds.setdefault((pack_site(s), m), len(conds)) for s, m in cs]
                    if a_idx > 255 or any(c > 254 for c in c_idx):
                        raise Unsupported("pool overflow")
                    w1 = 0
                    for j, c in enumerate(c_idx):
                        w1 |= c << (8 * j)
                    ops_words += [kind | (q << 4) | (a_idx << 16) | (len(cs) << 24), w1]
                n += len(rnd)
                cum.append(n)
            cum += [n] * (MAX_ROUNDS - len(cum))
            ev = [ops_start, len(rounds), len(writes), base_n] + cum
            for off, old, new in writes:
                ev += [pack_site(off), old | (new << 8)]
            ev += [0, 0] * (MAX_WRITES - len(writes))
            events_words += ev
            max_rounds = max(max_rounds, len(rounds))
            max_ops = max(max_ops, len(ops))
            stats.append((len(ops), len(rounds), len(writes)))
        anchors_words = [0] * len(anchors)
        for packed, i in anchors.items():
            anchors_words[i] = packed
        conds_words = [0] * (2 * len(conds))
        for (packed, m), i in conds.items():
            conds_words[2 * i] = packed
            conds_words[2 * i + 1] = m
    except Unsupported as e:
        info["reason"] = str(e)
        return header, info

    def s32(w):
        return w - (1 << 32) if w >= (1 << 31) else w

    hdr_len = 13
    events_off = hdr_len
    ops_off = events_off + len(events_words)
    anchors_off = ops_off + len(ops_words)
    conds_off = anchors_off + len(anchors_words)
    pa_off = conds_off + len(conds_words)
    header = [DEV_VERSION, 1, nproc, events_off, ops_off, len(ops_words) // 2, anchors_off, len(anchors_words),
              conds_off, len(conds_words) // 2, max_rounds, max_ops, pa_off]
    words = header + events_words + ops_words + anchors_words + conds_words + proc_anchor
    info.update({"supported": True, "n_ops": len(ops_words) // 2, "n_anchors": len(anchors_words),
                 "n_conds": len(conds_words) // 2, "max_rounds": max_rounds, "max_ops": max_ops,
                 "per_event": stats, "bytes": 4 * len(words), "proc_anchor": proc_anchor})
    return [s32(w) for w in words], info
