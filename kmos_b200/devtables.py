"""Per-event lane tables for the shared-memory CUDA kernel (SEC_DEVICE of the model blob), version 2.

1. Event flattening.  For a chosen (process, site) the generated Fortran executes a fixed *sequence* of
put_/take_ routines (local_smart; kmos/io/__init__.py:305-465, 2219-2409), each one ``replace_species``,
a flat list of guarded ``del_proc`` and an if-tree of ``get_species`` probes ending in ``add_proc``
leaves.  The whole event is flattened at export time into

    writes   the replace_species calls (site, old, new), in order
    ops      every del/add candidate of every action, in textual (= execution) order:
             DEL_IF (q, anchor)            guard avail_sites(q, anchor, 2) /= 0, read at execution time
             ADD    (q, anchor, conds...)  leaf of the if-tree; conds = species tests on its path

While an event runs, the species on its own action sites are known at compile time (its conditions fix
them before, its actions after), so if-tree probes of those sites are folded here: ops that cannot fire
are dropped and the remaining probes only touch sites the event does not modify -- every lane can
evaluate them up front, independent of the order of the event's own writes.

2. Compact avail-site storage.  The reference keeps avail_sites(nr_of_proc, volume, 2) (base.mpy:88).
Two processes registered on the same site type whose conditions contradict each other on some site can
never be available on the same cell at the same time, so
    * plane 2 (site -> position) is stored per *exclusivity class* (a clique of such processes) and cell:
      one uint16 entry (member << POS_BITS | position), instead of one entry per process and cell;
    * plane 1 (position -> site) is stored per *arena*: two processes of one class share ncells entries,
      one list growing from the left, the other from the right (their lengths can never sum above ncells).
For the RuO2 model this is 36x400x2x2 B = 57.6 KB -> ~19 KB per replica, i.e. 3x more replicas per SM.
The order of every process' list -- all that determine_procsite can observe -- is unchanged.

3. Rounds.  Bit-exact parity needs the add/del calls to hit each list in the reference order, nothing
more.  Ops are list-scheduled into rounds such that two ops sharing a resource (an arena, or a class
entry of the same anchor cell) keep their textual order in successive rounds; the ops of one round touch
disjoint memory and run one per lane.

Section layout (int32 words, offsets relative to the section start):
    [0] version=2 [1] supported [2] n_events [3] events_off [4] ops_off [5] n_ops [6] op_stride
    [7] offsets_off [8] n_offsets [9] procinfo_off [10] n_classes [11] n_arenas [12] max_rounds
    [13] max_ops_per_event [14] max_ncond [15] spare slots per arena (most adds of one event into one arena)
    events   EVENT_WORDS each: w0 = ops_start | n_rounds<<16 | n_writes<<20 | min_q<<24
                               w1,w2 = cumulative op count after each round (8 x u8), w3 = 0
                               w4..w7 = writes: off_id | n<<5 | old<<8 | new<<12
    ops      op_stride words each: header = kind | ncond<<1 | off_id<<4 | q<<9 | cls<<15 | member<<20 |
                               arena<<23 | dir<<29 ; then ncond words off_id | n<<5 | mask<<8
    offsets  1 word each: (dx&255) | (dy&255)<<8 | (dz&255)<<16
    procinfo 1 word per process: arena | dir<<6 | cls<<7 | member<<12 | anchor_n<<15
"""
import os

EVENT_WORDS = 8
MAX_ROUNDS = 8
MAX_WRITES = 4
MAX_COND = 4
MAX_OFFSETS = 32
MAX_CLASS_MEMBERS = 7
POS_BITS = 13
KIND_DEL_IF, KIND_ADD = 0, 1
DEV_VERSION = 2
WARP = 32
HEADER_WORDS = 16


class Unsupported(Exception):
    pass


def _add4(a, b):
    return [a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]]


def flatten_event(ir, proc_index):
    """-> (base_n, writes, ops) for process `proc_index` (0-based) of a local_smart model.

    writes: [(off4_abs, old, new)]; ops: [(kind, q, anchor_off4_abs, [(off4_abs, mask)...], group)] with
    group = 2*action (the action's guarded dels) or 2*action+1 (its if-tree adds).
    Offsets are relative to the selected site's cell, 4th component = absolute site type.
    """
    calls = ir["run_proc"][proc_index]
    rsite = ir["routine_site"]
    all_mask = (1 << len(ir["species"])) - 1
    first = calls[0]
    base_n = rsite[first[1]] - first[2][3]
    known = {}
    seq = []
    for _c, rname, off in calls:
        rb = [off[0], off[1], off[2], base_n + off[3]]
        if rb[3] != rsite[rname]:
            raise Unsupported("inconsistent site type in run_proc of process %d" % (proc_index + 1))
        seq.append((ir["routines"][rname], rb))
    for stmts, rb in seq:
        for st in stmts:
            if st[0] == "replace":
                known.setdefault(tuple(_add4(rb, st[1])), st[2])
    writes, ops = [], []
    action = [0]

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
                ops.append((KIND_DEL_IF, st[1], _add4(rb, st[2]), [], 2 * action[0]))
            elif k == "add":
                if isinstance(st[1], list) or st[3] is not None:
                    raise Unsupported("nli/otf add in local_smart table")
                ops.append((KIND_ADD, st[1], _add4(rb, st[2]), list(conds), 2 * action[0] + 1))
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
        action[0] += 1
    # An event may touch one site twice -- the reference turns A -> B into take_A (A -> empty) followed by put_B
    # (empty -> B), e.g. the predation step of examples/render_Lotka_Volterra_model.py.  The kernels apply an
    # event's lattice writes concurrently, one per lane, so such a pair becomes one write: expected species of
    # the first call (what replace_species checks first, base.mpy:1205), final species of the last.  The list
    # operations in between keep their order and their view of the intermediate species (`known` above).
    merged = []
    for site, old, new in writes:
        for w in merged:
            if w[0] == site:
                if w[2] != old:
                    raise Unsupported("replace_species chain on one site does not connect")
                w[2] = new
                break
        else:
            merged.append([site, old, new])
    return base_n, [tuple(w) for w in merged], ops


def process_conditions(ir):
    """Full condition list of every process relative to its anchor site, read off the touchup if-trees
    (kmos/io/__init__.py:2411-2443): {q: [(off4_rel, mask), ...]} with off4_rel[3] = absolute site type."""
    rsite = ir["routine_site"]
    all_mask = (1 << len(ir["species"])) - 1
    out = {}

    def walk(block, base_n, conds):
        for st in block:
            if st[0] == "add" and not isinstance(st[1], list):
                a = st[2]
                rel = [([c[0][0] - a[0], c[0][1] - a[1], c[0][2] - a[2], c[0][3]], c[1]) for c in conds]
                if st[1] in out and sorted(map(repr, out[st[1]])) != sorted(map(repr, rel)):
                    raise Unsupported("process %d has two different condition sets" % st[1])
                out[st[1]] = rel
            elif st[0] == "select":
                site = [st[1][0], st[1][1], st[1][2], base_n + st[1][3]]
                seen = 0
                for key, body in st[2]:
                    mask = (all_mask & ~seen) if key is None else (sum(1 << s for s in set(key)) & ~seen)
                    seen |= mask
                    walk(body, base_n, conds + [(site, mask)])
    for name, stmts in ir["routines"].items():
        if name.lower().startswith("touchup_") and name in rsite:
            walk(stmts, rsite[name], [])
    return out


def exclusive(ca, cb):
    """True if two condition lists (same anchor type) contradict each other on some site."""
    for sa, ma in ca:
        for sb, mb in cb:
            if sa == sb and (ma & mb) == 0:
                return True
    return False


def exclusivity_classes(ir, proc_anchor):
    """Smallest clique cover of the `mutually exclusive` relation -> (classes, cls_of, member_of).

    Every class costs one uint16 plane of ncells entries in shared memory, so the cover is searched exactly
    (depth first, first-fit order, bounded number of nodes; the first leaf is the greedy first-fit cover, so a
    cut-off search is never worse than that).  RuO2: 6 classes instead of first-fit's 7."""
    nproc = len(ir["procs"])
    conds = process_conditions(ir)

    def excl(a, b):
        return a in conds and b in conds and proc_anchor[a - 1] == proc_anchor[b - 1] and \
            exclusive(conds[a], conds[b])

    nodes = list(range(1, nproc + 1))
    best = [None]
    budget = [200000]

    def rec(i, classes):
        if best[0] is not None and len(classes) >= len(best[0]):
            return
        if i == len(nodes):
            best[0] = [list(c) for c in classes]
            return
        budget[0] -= 1
        if budget[0] < 0 and best[0] is not None:
            return
        q = nodes[i]
        for cl in classes:
            if len(cl) < MAX_CLASS_MEMBERS and all(excl(q, c) for c in cl):
                cl.append(q)
                rec(i + 1, classes)
                cl.pop()
        classes.append([q])
        rec(i + 1, classes)
        classes.pop()

    rec(0, [])
    classes = best[0]
    cls_of, member_of = {}, {}
    for ci, cl in enumerate(classes):
        for mi, q in enumerate(cl):
            cls_of[q] = ci
            member_of[q] = mi + 1
    return classes, cls_of, member_of


def schedule_rounds(ops, lists_of, entry_of, width=None):
    """List scheduling into rounds of at most WARP ops (one op per lane, rounds separated by __syncwarp).

    Ordering that must be kept from the reference's textual order:
      * ops on the same process list (``lists_of(op)``: its nr_of_sites counter and list) -- always;
      * on one class entry (``entry_of(op)``: exclusivity class x anchor cell) an ADD must follow every
        earlier guarded del of an earlier group (the del that frees the entry for it).
    Nothing else: at most one member of a class is registered on a cell, so
      - guarded dels of different processes on one entry: at most one fires, the others only read;
      - adds of one group (one action's if-tree) on one entry: at most one can fire;
      - a guarded del of q after an add of q' != q: if the add fired the entry was free before it and holds
        q' after it -- the del is a no-op either way; if it did not fire they do not interact;
      - two adds of different groups can both fire only with a del of the first process in between, which is
        ordered after the first add by its list and before the second add by the rule above.
    """
    rounds = []
    last_list = {}
    entry_dels = {}  # entry -> {group: last round of a guarded del}
    for i, op in enumerate(ops):
        kind, group = op[0], op[4]
        r = -1
        for x in lists_of(op):
            r = max(r, last_list.get(x, -1))
        ed = entry_dels.setdefault(entry_of(op), {})
        if kind == KIND_ADD:
            for g, rr in ed.items():
                if g < group:
                    r = max(r, rr)
        r += 1
        while True:
            if r == len(rounds):
                rounds.append([])
            if len(rounds[r]) < (width or WARP):
                break
            r += 1
        rounds[r].append(i)
        for x in lists_of(op):
            last_list[x] = r
        if kind == KIND_DEL_IF:
            ed[group] = max(ed.get(group, -1), r)
    return rounds


def compile_device_tables(ir, asm=None):
    info = {"supported": False}
    nproc = len(ir["procs"])
    header = [DEV_VERSION, 0] + [0] * (HEADER_WORDS - 2)
    if ir["backend"] == "lat_int":
        return compile_latint_tables(ir)
    if ir["backend"] != "local_smart":
        info["reason"] = "lane tables are generated for local_smart and lat_int only"
        return header, info
    try:
        from .tables import proc_anchor_types
        proc_anchor = proc_anchor_types(ir)
        if any(a == 0 for a in proc_anchor):
            raise Unsupported("a process is registered on several site types")
        if nproc > 64:
            raise Unsupported("more than 64 processes")
        if len(ir["species"]) > 16:
            raise Unsupported("more than 16 species")
        if ir["spuck"] > 7:
            raise Unsupported("more than 7 sites per cell")
        classes, cls_of, member_of = exclusivity_classes(ir, proc_anchor)
        if len(classes) > 32:
            raise Unsupported("more than 32 exclusivity classes")
        # arenas: two mutually exclusive processes share one (left list, right list); pairs are chosen by
        # a greedy maximum matching on the exclusivity graph (not restricted to one class)
        conds = process_conditions(ir)

        def can_pair(a, c):
            return (a != c and a in conds and c in conds and proc_anchor[a - 1] == proc_anchor[c - 1]
                    and exclusive(conds[a], conds[c]))

        def greedy_matching(order):
            match = {}
            for q in order:
                if q in match:
                    continue
                for c in order:
                    if c not in match and can_pair(q, c):
                        match[q], match[c] = c, q
                        break
            return match

        candidates = []
        # (a) neighbours inside each exclusivity class
        m = {}
        for cl in classes:
            for i in range(0, len(cl) - 1, 2):
                m[cl[i]], m[cl[i + 1]] = cl[i + 1], cl[i]
        candidates.append(m)
        # (b) greedy matchings: fewest possible partners first, then a few deterministic shuffles
        degree = {q: sum(1 for c in range(1, nproc + 1) if can_pair(q, c)) for q in range(1, nproc + 1)}
        candidates.append(greedy_matching(sorted(range(1, nproc + 1), key=lambda q: (degree[q], q))))
        import random
        rnd = random.Random(12345)
        for _ in range(32):
            order = list(range(1, nproc + 1))
            rnd.shuffle(order)
            candidates.append(greedy_matching(order))
        partner = max(candidates, key=len)
        arena_of, dir_of = {}, {}
        n_arenas = 0
        for q in range(1, nproc + 1):
            if q in arena_of:
                continue
            arena_of[q], dir_of[q] = n_arenas, 0
            if q in partner:
                arena_of[partner[q]], dir_of[partner[q]] = n_arenas, 1
            n_arenas += 1
        if n_arenas > 64:
            raise Unsupported("more than 64 arenas")

        offsets = {}

        def off_id(o):
            key = (o[0], o[1], o[2])
            for d in key:
                if not -128 <= d <= 127:
                    raise Unsupported("offset out of byte range")
            if key not in offsets:
                if len(offsets) == MAX_OFFSETS:
                    raise Unsupported("more than %d distinct neighbour offsets" % MAX_OFFSETS)
                offsets[key] = len(offsets)
            return offsets[key]

        off_id([0, 0, 0])
        flat = []
        max_ncond = 0
        for p in range(nproc):
            base_n, writes, ops = flatten_event(ir, p)
            if base_n != proc_anchor[p]:
                raise Unsupported("process %d is selected on site type %d but registered on %d"
                                  % (p + 1, base_n, proc_anchor[p]))
            for _k, q, aoff, cs, _g in ops:
                if aoff[3] != proc_anchor[q - 1]:
                    raise Unsupported("anchor site type mismatch")
                max_ncond = max(max_ncond, len(cs))
            if len(writes) > MAX_WRITES:
                raise Unsupported("event writes %d sites" % len(writes))
            flat.append((base_n, writes, ops))
        if max_ncond > MAX_COND:
            raise Unsupported("add with %d dynamic conditions" % max_ncond)
        op_stride = 1 + max_ncond

        ops_words, events_words, stats = [], [], []
        max_rounds = max_ops = 0
        spare = 0
        for _b, _w, ops in flat:
            per_arena = {}
            for kind, q, _a, _c, _g in ops:
                if kind == KIND_ADD:
                    per_arena[arena_of[q]] = per_arena.get(arena_of[q], 0) + 1
            spare = max([spare] + list(per_arena.values()))
        for p, (base_n, writes, ops) in enumerate(flat):
            rounds = schedule_rounds(ops, lambda op: [op[1]],
                                     lambda op: (cls_of[op[1]], op[2][0], op[2][1], op[2][2]))
            if len(rounds) > MAX_ROUNDS:
                raise Unsupported("event needs %d rounds" % len(rounds))
            ops_start = len(ops_words) // op_stride
            if ops_start >= 1 << 16:
                raise Unsupported("too many ops")
            cum, n = [], 0
            for rnd in rounds:
                for i in rnd:
                    kind, q, aoff, cs, _g = ops[i]
                    hdr = (kind | (len(cs) << 1) | (off_id(aoff) << 4) | ((q - 1) << 9) | (cls_of[q] << 15) |
                           (member_of[q] << 20) | (arena_of[q] << 23) | (dir_of[q] << 29))
                    words = [hdr]
                    for s, m in cs:
                        words.append(off_id(s) | (s[3] << 5) | (m << 8))
                    words += [0] * (op_stride - len(words))
                    ops_words += words
                n += len(rnd)
                if n > 255:
                    raise Unsupported("more than 255 ops in one event")
                cum.append(n)
            cum += [n] * (MAX_ROUNDS - len(cum))
            min_q = min([q for _k, q, _a, _c, _g in ops] + [nproc]) - 1
            w0 = ops_start | (len(rounds) << 16) | (len(writes) << 20) | (min_q << 24)
            w1 = cum[0] | (cum[1] << 8) | (cum[2] << 16) | (cum[3] << 24)
            w2 = cum[4] | (cum[5] << 8) | (cum[6] << 16) | (cum[7] << 24)
            ev = [w0, w1, w2, 0]
            for off, old, new in writes:
                ev.append(off_id(off) | (off[3] << 5) | (old << 8) | (new << 12))
            ev += [0] * (EVENT_WORDS - len(ev))
            events_words += ev
            max_rounds = max(max_rounds, len(rounds))
            max_ops = max(max_ops, len(ops))
            stats.append((len(ops), len(rounds), len(writes)))
        offsets_words = [0] * len(offsets)
        for (dx, dy, dz), i in offsets.items():
            offsets_words[i] = (dx & 255) | ((dy & 255) << 8) | ((dz & 255) << 16)
        procinfo = [arena_of[q] | (dir_of[q] << 6) | (cls_of[q] << 7) | (member_of[q] << 12) |
                    (proc_anchor[q - 1] << 15) for q in range(1, nproc + 1)]
    except Unsupported as e:
        info["reason"] = str(e)
        return header, info

    def s32(w):
        return w - (1 << 32) if w >= (1 << 31) else w

    events_off = HEADER_WORDS
    ops_off = events_off + len(events_words)
    offsets_off = ops_off + len(ops_words)
    procinfo_off = offsets_off + len(offsets_words)
    header = [DEV_VERSION, 1, nproc, events_off, ops_off, len(ops_words) // op_stride, op_stride, offsets_off,
              len(offsets_words), procinfo_off, len(classes), n_arenas, max_rounds, max_ops, max_ncond, spare]
    words = header + events_words + ops_words + offsets_words + procinfo
    info.update({"supported": True, "n_ops": len(ops_words) // op_stride, "op_stride": op_stride,
                 "n_offsets": len(offsets_words), "n_classes": len(classes), "n_arenas": n_arenas,
                 "classes": classes, "spare": spare, "max_rounds": max_rounds, "max_ops": max_ops, "max_ncond": max_ncond,
                 "per_event": stats, "bytes": 4 * len(words), "proc_anchor": proc_anchor})
    return [s32(w) for w in words], info


# ======================================================================================================
# lat_int: tables for the warp-per-replica kernel (kb_latint.cuh)
# ======================================================================================================
#
# run_proc_<group>(cell) of the lat_int generator (kmos/io/__init__.py:1793-1983) has a fixed shape:
#     del_proc(nli_g(cell + o), cell + o + (0,0,0,1))   for a sorted list of (g, o)
#     replace_species(...)                               the group's actions
#     add_proc(nli_g(cell + o), cell + o + (0,0,0,1))   the same list
# and nli_<g>(cell) (io/__init__.py:1985-2057) is a decision tree over lattice species ending in a process
# number or 0.  One lane evaluates one (g, o) pair; lanes that resolve to the same process run in lane
# (= textual) order, everything else concurrently.
#
# Section layout (int32 words, offsets relative to the section start):
#     [0] version=3 [1] supported [2] n_events [3] events_off [4] ops_off [5] n_ops [6] nodes_off
#     [7] offsets_off [8] n_offsets [9] n_nodes [10] n_species [11] writes_off [12] n_writes
#     [13] max_ops_per_phase [14..15] reserved
#     events  4 words per process: dels_start, n_dels | n_adds<<16, adds_start, writes_start | n_writes<<16
#     ops     1 word: root node | off_id<<16                      (off_id: cell the function is evaluated on)
#     writes  1 word: off_id | n<<8 | (old+1)<<16 | (new+1)<<24   (species + 1: null_species = 0)
#     nodes   (1 + n_species) words: off_id | n<<8 ; per species: 0x80000000 | process  or  child node index
#     offsets 1 word: (dx&255) | (dy&255)<<8 | (dz&255)<<16
LATINT_VERSION = 3


def _tree_nodes(stmts, n_species, off_id, nodes, memo):
    """Flatten the statement list of an nli function into decision-tree nodes; returns the reference of its
    root: 0x80000000 | process for a leaf, else a node index."""
    key = repr(stmts)
    if key in memo:
        return memo[key]
    if not stmts:
        return 0x80000000
    st = stmts[0]
    if st[0] == "return":
        return 0x80000000 | int(st[1])
    if st[0] != "select":
        raise Unsupported("statement %r in nli function" % st[0])
    rest = stmts[1:]
    idx = len(nodes)
    nodes.append(None)
    children = []
    for s in range(n_species):
        body = None
        default = None
        for keyset, b in st[2]:
            if keyset is None:
                default = b
            elif s in keyset and body is None:
                body = b
        chosen = body if body is not None else default
        seq = (chosen if chosen is not None else []) + rest
        children.append(_tree_nodes(seq, n_species, off_id, nodes, memo))
    nodes[idx] = [off_id(st[1]) | (st[1][3] << 8)] + children
    memo[key] = idx
    return idx


def compile_latint_tables(ir):
    info = {"supported": False}
    header = [LATINT_VERSION, 0] + [0] * (HEADER_WORDS - 2)
    if ir["backend"] != "lat_int":
        info["reason"] = "not a lat_int model"
        return header, info
    nproc = len(ir["procs"])
    n_species = len(ir["species"])
    try:
        if nproc > 256:
            raise Unsupported("more than 256 processes")
        offsets = {}

        def off_id(o):
            key = (o[0], o[1], o[2])
            for d in key:
                if not -128 <= d <= 127:
                    raise Unsupported("offset out of byte range")
            if key not in offsets:
                if len(offsets) == 255:
                    raise Unsupported("too many distinct offsets")
                offsets[key] = len(offsets)
            return offsets[key]

        off_id([0, 0, 0])
        nodes, roots = [], {}
        for name, stmts in sorted(ir["nli"].items()):
            root = _tree_nodes(stmts, n_species, off_id, nodes, {})
            if root & 0x80000000:  # a constant function: wrap it into a one-node tree on the cell itself
                idx = len(nodes)
                nodes.append([off_id([0, 0, 0]) | (1 << 8)] + [root] * n_species)
                root = idx
            roots[name] = root
        if len(nodes) >= 1 << 16:
            raise Unsupported("decision trees too large")
        ops_words, writes_words, events_words = [], [], []
        cache = {}
        max_phase = 0
        for p in range(nproc):
            calls = ir["run_proc"][p]
            if len(calls) != 1 or calls[0][2] != [0, 0, 0, -1]:
                raise Unsupported("run_proc_nr of process %d is not a single cell routine" % (p + 1))
            rname = calls[0][1]
            if rname not in cache:
                dels, adds, writes = [], [], []
                phase = 0
                for st in ir["routines"][rname]:
                    if st[0] in ("del", "add"):
                        if not isinstance(st[1], list) or st[1][0] != "nli":
                            raise Unsupported("lat_int op without nli function")
                        celloff, siteoff = st[1][2], st[2]
                        if celloff[3] != 0 or siteoff[:3] != celloff[:3] or siteoff[3] != 1:
                            raise Unsupported("lat_int op not registered on site 1 of the evaluated cell")
                        word = roots[st[1][1]] | (off_id(celloff) << 16)
                        if st[0] == "del":
                            if phase != 0:
                                raise Unsupported("del after lattice update")
                            dels.append(word)
                        else:
                            phase = 2
                            adds.append(word)
                    elif st[0] == "replace":
                        if phase == 2:
                            raise Unsupported("lattice update after add")
                        phase = 1
                        writes.append(off_id(st[1]) | (st[1][3] << 8) | ((st[2] + 1) << 16) | ((st[3] + 1) << 24))
                    else:
                        raise Unsupported("statement %r in lat_int run_proc" % st[0])
                ev = [len(ops_words), len(dels) | (len(adds) << 16), len(ops_words) + len(dels),
                      len(writes_words) | (len(writes) << 16)]
                ops_words += dels + adds
                writes_words += writes
                max_phase = max(max_phase, len(dels), len(adds))
                cache[rname] = ev
            events_words += cache[rname]
        offsets_words = [0] * len(offsets)
        for (dx, dy, dz), i in offsets.items():
            offsets_words[i] = (dx & 255) | ((dy & 255) << 8) | ((dz & 255) << 16)
        nodes_words = [w for nd in nodes for w in nd]
    except Unsupported as e:
        info["reason"] = str(e)
        return header, info

    def s32(w):
        return w - (1 << 32) if w >= (1 << 31) else w

    events_off = HEADER_WORDS
    ops_off = events_off + len(events_words)
    writes_off = ops_off + len(ops_words)
    nodes_off = writes_off + len(writes_words)
    offsets_off = nodes_off + len(nodes_words)
    header = [LATINT_VERSION, 1, nproc, events_off, ops_off, len(ops_words), nodes_off, offsets_off,
              len(offsets_words), len(nodes), n_species, writes_off, len(writes_words), max_phase, 0, 0]
    words = header + events_words + ops_words + writes_words + nodes_words + offsets_words
    info.update({"supported": True, "n_ops": len(ops_words), "n_nodes": len(nodes), "n_offsets": len(offsets_words),
                 "max_ops_per_phase": max_phase, "bytes": 4 * len(words)})
    return [s32(w) for w in words], info


# ======================================================================================================
# otf: lane tables for the production kernel (kb_otf_fast.cuh)
# ======================================================================================================
#
# run_proc_<proc>(cell) of the otf generator (kmos/io/__init__.py:3328-3596) has a fixed shape:
#     if (can_do(q, cell + o)) del_proc(q, cell + o)                                   disabled processes
#     replace_species(...)                                                             the actions
#     if (can_do(q, cell + o)) update_rates_matrix(q, cell + o, gr_q(cell + o'))       changed bystanders
#     add_proc(q, cell + o, gr_q(cell + o')) / select case nests of them               enabled processes
# Executed by one lane every statement is a chain of dependent DRAM accesses.  The first three blocks go to
# the lanes of the warp, one statement per lane (statements on one process keep their textual order, as in
# the lat_int kernel); the last block stays byte-code -- a routine of its own, run by lane 0.
#
# Section layout (int32 words, offsets relative to the section start):
#     [0] version=6 [1] supported [2] n_events [3] events_off [4] ops_off [5] n_ops [6] conds_off [7] n_conds
#     events  8 words per process: dels_start, n_dels, writes_start, n_writes, upds_start, n_upds,
#             tail routine id (-1: none), n_adds (ops behind the updates)
#     ops     12 words: del     q, dx, dy, dz, n, 0...
#                       write   old, dx, dy, dz, n, new, 0...
#                       update  q, dx, dy, dz, n, gr id, gx, gy, gz, gn, 0, 0     (gr evaluated on cell + g)
#                       add     q, dx, dy, dz, n, gr id, gx, gy, gz, gn, conds_start, n_conds
#     conds   5 words: dx, dy, dz, n, species mask (bit 31: the null species passes)
# The add block is the generator's if-tree (select case nests, io/__init__.py:2568-2655) flattened: every
# add_proc statement with the case labels on its path as conditions, in textual order.  The lattice does not
# change while the block runs and exactly one case of a select is taken, so the flattened list appends the same
# processes in the same order.  Blocks with more than 255 statements or 8 conditions stay a byte-code routine.
OTF_VERSION = 6
OTF_OP_WORDS = 12
OTF_MAX_CONDS = 8


def compile_otf_tables(ir, asm):
    """Lane tables of an otf model; assembles one tail routine per distinct run_proc routine into `asm`."""
    info = {"supported": False}
    header = [OTF_VERSION, 0] + [0] * (HEADER_WORDS - 2)
    nproc = len(ir["procs"])

    def only_adds(block):
        for st in block:
            if st[0] == "select":
                if not all(only_adds(body) for _k, body in st[2]):
                    return False
            elif not (st[0] == "add" and not isinstance(st[1], list) and st[3] is not None):
                return False
        return True

    n_species = len(ir["species"])

    def flatten_adds(block, path, out):
        for st in block:
            if st[0] == "add":
                out.append((st, list(path)))
            else:  # select: first matching label wins, `case default` takes what no label names (and null)
                seen = 0
                default_body = None
                for key, body in st[2]:
                    if key is None:
                        default_body = body
                        continue
                    mask = 0
                    for sp in key:
                        mask |= 1 << sp
                    mask &= ~seen
                    seen |= mask
                    flatten_adds(body, path + [st[1] + [mask]], out)
                if default_body is not None:
                    rest = ((1 << n_species) - 1) & ~seen
                    flatten_adds(default_body, path + [st[1] + [rest | (1 << 31)]], out)

    try:
        ops, events, cache, conds = [], [], {}, []
        for p in range(nproc):
            calls = ir["run_proc"][p]
            if len(calls) != 1 or calls[0][0] != "call" or calls[0][2] != [0, 0, 0, -1]:
                raise Unsupported("run_proc_nr of process %d is not a single cell routine" % (p + 1))
            rname = calls[0][1]
            if rname not in cache:
                dels, writes, upds = [], [], []
                stmts = list(ir["routines"][rname])
                i = 0
                while i < len(stmts) and stmts[i][0] == "if_can" and len(stmts[i][3]) == 1 and stmts[i][3][0][0] == "del":
                    st, d = stmts[i], stmts[i][3][0]
                    if isinstance(d[1], list) or d[1] != st[1] or d[2] != st[2]:
                        raise Unsupported("guarded del of another process or site")
                    dels.append([st[1]] + st[2] + [0] * 7)
                    i += 1
                while i < len(stmts) and stmts[i][0] == "replace":
                    st = stmts[i]
                    writes.append([st[2]] + st[1] + [st[3]] + [0] * 6)
                    i += 1
                if len(set(tuple(w[1:5]) for w in writes)) != len(writes):
                    raise Unsupported("%s: two lattice writes on one site" % rname)  # lanes would race
                while i < len(stmts) and stmts[i][0] == "if_can" and len(stmts[i][3]) == 1 and \
                        stmts[i][3][0][0] == "update_rate":
                    st, u = stmts[i], stmts[i][3][0]
                    if u[1] != st[1] or u[2] != st[2] or u[3][0] != "gr":
                        raise Unsupported("guarded update of another process or site")
                    upds.append([st[1]] + st[2] + [asm.gr_id(u[3][1])] + u[3][2] + [0, 0])
                    i += 1
                tail = stmts[i:]
                if not only_adds(tail):
                    raise Unsupported("%s: statements after the update block other than add_proc" % rname)
                if len(dels) > 255 or len(upds) > 255 or len(writes) > 32:
                    raise Unsupported("%s: too many statements" % rname)
                adds, flat = [], []
                flatten_adds(tail, [], flat)
                if flat and len(flat) <= 255 and all(len(path) <= OTF_MAX_CONDS for _st, path in flat) and \
                        os.environ.get("KMOS_B200_OTF_FLATTEN", "1") != "0":   # "0": keep the routine (tests)
                    for st, path in flat:
                        adds.append([st[1]] + st[2] + [asm.gr_id(st[3][1])] + st[3][2] + [len(conds), len(path)])
                        conds += path
                    tail = []
                tail_id = asm.anon_routine("__otf_tail_" + rname, tail) if tail else -1
                ev = [len(ops), len(dels)]
                ops += dels
                ev += [len(ops), len(writes)]
                ops += writes
                ev += [len(ops), len(upds), tail_id, len(adds)]
                ops += upds + adds
                cache[rname] = ev
            events += cache[rname]
    except Unsupported as e:
        info["reason"] = str(e)
        return header, info
    events_off = HEADER_WORDS
    ops_off = events_off + len(events)
    conds_off = ops_off + OTF_OP_WORDS * len(ops)
    header = [OTF_VERSION, 1, nproc, events_off, ops_off, len(ops), conds_off, len(conds)] + [0] * (HEADER_WORDS - 8)
    assert all(len(op) == OTF_OP_WORDS for op in ops) and all(len(c) == 5 for c in conds)
    words = header + events + [w for op in ops for w in op] + [w for c in conds for w in c]

    def s32(w):
        return w - (1 << 32) if w >= (1 << 31) else w

    info.update({"supported": True, "n_ops": len(ops), "n_conds": len(conds), "bytes": 4 * len(words),
                 "routine_tails": sum(1 for ev in cache.values() if ev[6] >= 0)})
    return [s32(w) for w in words], info


# ======================================================================================================
# local_smart on lattices that do not fit shared memory: tables for the warp-per-replica HBM kernel
# ======================================================================================================
#
# Same flattened events as the shared-memory tables (flatten_event) but kept in textual order and executed on
# the canonical per-process planes: the only ordering that matters is between ops of the same process, which
# the kernel resolves with a warp match, like for lat_int.
#
# Section layout (int32 words):
#     [0] version=4 [1] supported [2] n_events [3] events_off [4] ops_off [5] n_ops [6] op_stride
#     [7] offsets_off [8] n_offsets [9..10] 0 [11] writes_off [12] n_writes [13] max_ops_per_event
#     events  4 words per process: ops_start (op index), n_ops, writes_start | n_writes<<16, base site type
#     ops     op_stride words: kind | ncond<<1 | off_id<<4 | q<<12 ; then ncond words off_id | n<<8 | mask<<16
#     writes  off_id | n<<8 | (old+1)<<16 | (new+1)<<24
#     offsets (dx&255) | (dy&255)<<8 | (dz&255)<<16
HBM_VERSION = 4


def compile_hbm_tables(ir):
    info = {"supported": False}
    header = [HBM_VERSION, 0] + [0] * (HEADER_WORDS - 2)
    if ir["backend"] != "local_smart":
        info["reason"] = "not a local_smart model"
        return header, info
    nproc = len(ir["procs"])
    try:
        from .tables import proc_anchor_types
        proc_anchor = proc_anchor_types(ir)
        if any(a == 0 for a in proc_anchor):
            raise Unsupported("a process is registered on several site types")
        if nproc > 256:
            raise Unsupported("more than 256 processes")
        if len(ir["species"]) > 16:
            raise Unsupported("more than 16 species")
        offsets = {}

        def off_id(o):
            key = (o[0], o[1], o[2])
            for d in key:
                if not -128 <= d <= 127:
                    raise Unsupported("offset out of byte range")
            if key not in offsets:
                if len(offsets) == 255:
                    raise Unsupported("too many distinct offsets")
                offsets[key] = len(offsets)
            return offsets[key]

        off_id([0, 0, 0])
        flat = [flatten_event(ir, p) for p in range(nproc)]
        max_ncond = max([len(op[3]) for _b, _w, ops in flat for op in ops] + [0])
        stride = 1 + max_ncond
        ops_words, writes_words, events_words = [], [], []
        max_ops = 0
        for p, (base_n, writes, ops) in enumerate(flat):
            if base_n != proc_anchor[p]:
                raise Unsupported("process %d is selected on another site type than it is registered on" % (p + 1))
            ev = [len(ops_words) // stride, len(ops), len(writes_words) | (len(writes) << 16), base_n]
            for kind, q, aoff, cs, _g in ops:
                if aoff[3] != proc_anchor[q - 1]:
                    raise Unsupported("anchor site type mismatch")
                words = [kind | (len(cs) << 1) | (off_id(aoff) << 4) | (q << 12)]
                for site, mask in cs:
                    words.append(off_id(site) | (site[3] << 8) | (mask << 16))
                words += [0] * (stride - len(words))
                ops_words += words
            for off, old, new in writes:
                writes_words.append(off_id(off) | (off[3] << 8) | ((old + 1) << 16) | ((new + 1) << 24))
            events_words += ev
            max_ops = max(max_ops, len(ops))
        offsets_words = [0] * len(offsets)
        for (dx, dy, dz), i in offsets.items():
            offsets_words[i] = (dx & 255) | ((dy & 255) << 8) | ((dz & 255) << 16)
    except Unsupported as e:
        info["reason"] = str(e)
        return header, info

    def s32(w):
        return w - (1 << 32) if w >= (1 << 31) else w

    events_off = HEADER_WORDS
    ops_off = events_off + len(events_words)
    writes_off = ops_off + len(ops_words)
    offsets_off = writes_off + len(writes_words)
    header = [HBM_VERSION, 1, nproc, events_off, ops_off, len(ops_words) // stride, stride, offsets_off,
              len(offsets_words), 0, 0, writes_off, len(writes_words), max_ops, 0, 0]
    words = header + events_words + ops_words + writes_words + offsets_words
    info.update({"supported": True, "n_ops": len(ops_words) // stride, "op_stride": stride,
                 "max_ops_per_event": max_ops, "bytes": 4 * len(words)})
    return [s32(w) for w in words], info
