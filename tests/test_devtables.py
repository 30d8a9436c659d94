"""The per-event lane tables (kmos_b200/devtables.py) must reproduce the reference's avail_sites ORDER.

A tiny Python model of the CUDA kernel's event phase (writes, then rounds of concurrent per-process
list operations) is driven by the (proc, site) sequence of the oracle; after every event lattice,
nr_of_sites and both planes of avail_sites must equal the oracle's, which executes the generated
statements one by one in textual order.
"""
import numpy as np
import pytest

from conftest import load_model
from kmos_b200 import devtables as dt
from oracle import oracle


def _s8(b):
    return b - 256 if b >= 128 else b


class EventModel(object):
    """Python model of the kernel's compact storage and event phase (kb_smem.cuh).  The ops of one round
    run concurrently on the GPU, so here they are executed in a random order: the result must not depend
    on it."""

    def __init__(self, blob, info, o, spare=None):
        from kmos_b200.tables import SEC_DEVICE
        nsec = blob[13]
        for i in range(nsec):
            if blob[14 + 3 * i] == SEC_DEVICE:
                off, ln = blob[14 + 3 * i + 1], blob[14 + 3 * i + 2]
        self.d = [int(x) & 0xFFFFFFFF for x in blob[off:off + ln]]
        d = self.d
        assert d[0] == dt.DEV_VERSION and d[1] == 1
        self.events_off, self.ops_off, self.stride = d[3], d[4], d[6]
        self.offsets = [(_s8(w & 255), _s8((w >> 8) & 255), _s8((w >> 16) & 255)) for w in d[d[7]:d[7] + d[8]]]
        self.procinfo = d[d[9]:d[9] + d[2]]
        self.n_classes, self.n_arenas = d[10], d[11]
        self.size = o.size
        self.spuck = o.spuck
        self.P = o.n_proc
        self.C = o.volume // o.spuck
        self.cap = self.C + (d[15] if spare is None else spare)
        self.rng = np.random.RandomState(5)
        self.lattice = o.lattice.copy()
        self.n = [int(x) for x in o.nr_of_sites]
        self.p1 = [None] * (self.n_arenas * self.cap)
        self.p2 = [0] * (self.n_classes * self.C)
        av = o.avail_sites
        for q in range(self.P):
            arena, dirn, cls, member, an = self.pinfo(q)
            for k in range(self.n[q]):
                cell = (int(av[q, k, 0]) - 1) // self.spuck
                assert (int(av[q, k, 0]) - 1) % self.spuck + 1 == an
                self.p1[self.slot(arena, dirn, k)] = cell
                assert self.p2[cls * self.C + cell] == 0, "two members of a class on one cell"
                self.p2[cls * self.C + cell] = (member << dt.POS_BITS) | (k + 1)

    def pinfo(self, q):
        w = self.procinfo[q]
        return w & 63, (w >> 6) & 1, (w >> 7) & 31, (w >> 12) & 7, (w >> 15) & 7

    def slot(self, arena, dirn, k):
        return arena * self.cap + (self.cap - 1 - k if dirn else k)

    def cell(self, x, y, z):
        L = self.size
        return (x % L[0]) + L[0] * ((y % L[1]) + L[1] * (z % L[2]))

    def run(self, proc, site):
        d = self.d
        c = (site - 1) // self.spuck
        x, y, z = c % self.size[0], (c // self.size[0]) % self.size[1], c // (self.size[0] * self.size[1])
        nb = [self.cell(x + o[0], y + o[1], z + o[2]) for o in self.offsets]
        ev = self.events_off + (proc - 1) * dt.EVENT_WORDS
        w0, w1, w2 = d[ev], d[ev + 1], d[ev + 2]
        ops_start, n_rounds, n_writes = w0 & 0xFFFF, (w0 >> 16) & 15, (w0 >> 20) & 15
        ends = w1 | (w2 << 32)
        pre = self.lattice.copy()
        for w in range(n_writes):
            ww = d[ev + 4 + w]
            idx = nb[ww & 31] * self.spuck + ((ww >> 5) & 7) - 1
            assert self.lattice[idx] == (ww >> 8) & 15
            self.lattice[idx] = (ww >> 12) & 15
        start = 0
        for r in range(n_rounds):
            endr = (ends >> (8 * r)) & 255
            assert endr - start <= 32
            order = list(range(start, endr))
            self.rng.shuffle(order)
            for i in order:
                base = self.ops_off + (ops_start + i) * self.stride
                h = d[base]
                kind, ncond, ca = h & 1, (h >> 1) & 7, nb[(h >> 4) & 31]
                q, cls, member = (h >> 9) & 63, (h >> 15) & 31, (h >> 20) & 7
                arena, dirn = (h >> 23) & 63, (h >> 29) & 1
                assert (arena, dirn, cls, member) == self.pinfo(q)[:4]
                if kind == dt.KIND_ADD:
                    ok = True
                    for j in range(ncond):
                        cw = d[base + 1 + j]
                        idx = nb[cw & 31] * self.spuck + ((cw >> 5) & 7) - 1
                        assert pre[idx] == self.lattice[idx], "probe of a site the event writes"
                        ok = ok and (((cw >> 8) >> self.lattice[idx]) & 1)
                    if ok:
                        assert self.p2[cls * self.C + ca] == 0, "class entry occupied"
                        s = self.slot(arena, dirn, self.n[q])
                        self.p1[s] = ca
                        self.p2[cls * self.C + ca] = (member << dt.POS_BITS) | (self.n[q] + 1)
                        self.n[q] += 1
                else:
                    e = self.p2[cls * self.C + ca]
                    if (e >> dt.POS_BITS) == member:
                        pos = e & ((1 << dt.POS_BITS) - 1)
                        nq = self.n[q]
                        last = self.p1[self.slot(arena, dirn, nq - 1)]
                        if pos < nq:
                            self.p1[self.slot(arena, dirn, pos - 1)] = last
                            self.p2[cls * self.C + last] = (member << dt.POS_BITS) | pos
                        self.p1[self.slot(arena, dirn, nq - 1)] = None
                        self.p2[cls * self.C + ca] = 0
                        self.n[q] = nq - 1
            start = endr

    def check(self, o):
        assert np.array_equal(self.lattice, o.lattice)
        n = o.nr_of_sites
        av = o.avail_sites
        assert list(n) == self.n
        for q in range(self.P):
            arena, dirn, cls, member, an = self.pinfo(q)
            mine = [self.p1[self.slot(arena, dirn, k)] * self.spuck + an for k in range(self.n[q])]
            assert mine == list(av[q, :n[q], 0]), "avail_sites order differs for process %d" % (q + 1)
            for k, s in enumerate(mine):
                assert self.p2[cls * self.C + (s - 1) // self.spuck] == (member << dt.POS_BITS) | (k + 1)
        # every non-empty class entry belongs to exactly one registered (process, cell)
        assert sum(1 for e in self.p2 if e) == sum(self.n)


@pytest.mark.parametrize("name,size,steps", [
    ("ab_local_smart", [20, 20], 3000),
    ("mini_101_local_smart", [5, 4], 300),
    ("zgb_local_smart", [12, 10], 3000),
    ("ruo2_local_smart", [6, 5], 4000),
    ("mini_101_local_smart", [3, 3], 2000),
    ("ruo2_local_smart", [20, 20], 1500),
    ("pairwise_local_smart", [8, 8], 2000),
    ("pt111_local_smart", [7, 6], 2000), ("einsd_local_smart", [19], 1500),
    ("multidentate_local_smart", [9, 8], 2000),
])
def test_event_tables_reproduce_avail_order(name, size, steps):
    ir, blob, info = load_model(name)
    assert info["device"]["supported"], info["device"].get("reason")
    rng = np.random.RandomState(7)
    rates = np.exp(rng.uniform(-1.5, 1.5, len(ir["procs"])))
    o = oracle.Oracle(blob, size, seed=11, rates=rates)
    m = EventModel(blob, info, o)
    m.check(o)
    for i in range(steps):
        p, s, st = _next(o)
        assert st == oracle.OK
        m.run(p, s)
        if i % 97 == 0 or i == steps - 1:
            m.check(o)


def _next(o):
    """One oracle event through get_next_kmc_step + run_proc_nr, returning the (proc, site) it executed.
    That pair of calls does not advance kmc_step (the Philox counter), so the key is changed per event."""
    o.L.kmos_oracle_seed(o.h, oracle.RNG_PHILOX, int(o.procstat.sum()) * 7919 + 13, 0)
    p, s, st = o.get_next_kmc_step()
    if st == oracle.OK:
        o.run_proc_nr(p, s)
    return p, s, st


def test_lane_group_width_model():
    """codegen: rounds per lane-group width, the lock-step expectation and the score the engine ranks widths by
    (measured best widths on B200: RuO2 20x20 8 for a full batch and 16 for a 2048-replica shard, ZGB 32x32 and
    pairwise 30x30 16, AB and mini_101 20x20 8 -- DESIGN.md 4.1)."""
    from kmos_b200 import codegen
    assert codegen.expected_max_rounds([3, 3, 3], 4) == 3.0
    assert abs(codegen.expected_max_rounds([1, 3], 2) - 2.5) < 1e-12        # P(max = 1) = 1/4
    assert codegen.expected_max_rounds([], 2) == 0.0
    # resident replicas per SM at 4 / 8 / 16 / 32 lanes per replica (None: width not offered), measured best
    measured_best = {"ruo2": ((None, 32, 32, 24), (8,)), "zgb": ((None, 24, 26, 24), (16,)),
                     "pairwise": ((None, 24, 24, 24), (16,)),
                     "ab": ((72, 64, 40, 24), (4, 8)),          # 5.99e9 and 5.97e9: a tie
                     "mini_101": ((80, 64, 40, 24), (4,)),      # 8.9e9 against 7.2e9 at 8 lanes
                     "ruo2 ": ((None, 14, 14, 14), (16,))}      # a 2048-replica shard: 14 replicas per SM at any width
    assert codegen.lane_group_widths(2) == [4, 8, 16, 32] and codegen.lane_group_widths(36) == [8, 16, 32]
    with pytest.raises(dt.Unsupported):
        codegen.analyse(load_model("ruo2_local_smart")[0], 4)
    for name, (resident, best) in measured_best.items():
        ir, _blob, _info = load_model(name.strip() + "_local_smart")
        an = codegen._flatten(ir)
        scores = {}
        for w, reps in zip((4, 8, 16, 32), resident):
            if reps is None:
                continue
            rounds = [len(r) for r in codegen._schedule(an, w)]
            assert max(len(x) for r in codegen._schedule(an, w) for x in r) <= w
            scores[w] = codegen.lane_group_score(an["nproc"], codegen.expected_max_rounds(rounds, 32 // w), 32 // w, reps)
        assert max(scores, key=scores.get) in best, (name, scores)
    # the exclusivity classes are a minimum clique cover: RuO2 needs 6 planes, not first-fit's 7
    ir, _b, _i = load_model("ruo2_local_smart")
    assert len(codegen.analyse(ir)["classes"]) == 6


def test_two_replace_species_calls_on_one_site_become_one_write():
    """examples/render_Lotka_Volterra_model.py, AB_reaction*: the reference turns A -> B on one site into take_A
    (A -> empty) followed by put_B (empty -> B).  The kernels apply an event's lattice writes concurrently, one per
    lane, so the pair must reach them as one write (first expected species, last new species); the list
    operations in between keep the intermediate species in view (found on the GPU in round 2: every replica
    stopped with a species-mismatch status)."""
    ir, _blob, _info = load_model("lotka_local_smart")
    A, B = ir["species"].index("A"), ir["species"].index("B")
    for p, name in enumerate(ir["procs"]):
        _base_n, writes, ops = dt.flatten_event(ir, p)
        sites = [tuple(w[0]) for w in writes]
        assert len(set(sites)) == len(sites), name
        if name.startswith("AB_reaction"):
            assert [(old, new) for _s, old, new in writes] == [(A, B)], (name, writes)
            assert len(ops) > 0


@pytest.mark.parametrize("name", ["pairwise_otf_otf", "intzgb_otf", "ruo2default_otf", "multidentate_otf", "hop3d_otf",
                                  "ab_otf", "mini_101_otf", "zgb_otf", "pt111_otf", "einsd_otf"])
def test_otf_lane_tables_partition_every_event_routine(name):
    """compile_otf_tables: every statement of run_proc_<proc> ends up in exactly one of the del / write / update
    blocks or in the tail routine, in order, with the operands of the statement."""
    from kmos_b200 import tables
    ir, _blob, info = load_model(name)
    assert info["device"]["supported"], info["device"].get("reason")
    blob, info2 = tables.build_blob(ir)
    words = None
    n_sections = int(blob[13])
    for i in range(n_sections):
        sid, off, ln = (int(v) for v in blob[14 + 3 * i:17 + 3 * i])
        if sid == tables.SEC_DEVICE:
            words = [int(w) for w in blob[off:off + ln]]
    assert words is not None and words[0] == dt.OTF_VERSION and words[1] == 1 and words[2] == len(ir["procs"])
    gr_by_id = {}
    for st_name, g in info2["gr"].items():
        gr_by_id[st_name] = g
    for p in range(len(ir["procs"])):
        ev = words[words[3] + 8 * p: words[3] + 8 * p + 8]
        stmts = ir["routines"][ir["run_proc"][p][0][1]]
        W = dt.OTF_OP_WORDS
        ops = lambda start, n: [words[words[4] + W * (start + i): words[4] + W * (start + i) + W] for i in range(n)]
        dels, writes, upds = ops(ev[0], ev[1]), ops(ev[2], ev[3]), ops(ev[4], ev[5])
        k = 0
        for op in dels:
            assert stmts[k][0] == "if_can" and [stmts[k][1]] + stmts[k][2] == op[:5] and stmts[k][3][0][0] == "del"
            k += 1
        for op in writes:
            assert stmts[k][0] == "replace" and [stmts[k][2]] + stmts[k][1] + [stmts[k][3]] == op[:6]
            k += 1
        for op in upds:
            u = stmts[k][3][0]
            assert stmts[k][0] == "if_can" and u[0] == "update_rate" and [stmts[k][1]] + stmts[k][2] == op[:5]
            assert op[6:10] == u[3][2]
            k += 1
        tail = stmts[k:]
        assert all(st[0] in ("add", "select") for st in tail)
        if ev[7]:   # the if-tree flattened: every add_proc statement with the case labels on its path
            assert ev[6] == -1

            def leaves(block, path):
                for st in block:
                    if st[0] == "add":
                        yield st, path
                    else:
                        for key, body in st[2]:
                            yield from leaves(body, path + [(st[1], key, st[2])])
            flat = list(leaves(tail, []))
            assert ev[7] == len(flat)
            n_species = len(ir["species"])
            for op, (st, path) in zip(ops(ev[4] + ev[5], ev[7]), flat):
                assert [st[1]] + st[2] == op[:5] and op[6:10] == st[3][2] and op[11] == len(path)
                for c, (site, key, cases) in enumerate(path):
                    cd = words[words[6] + 5 * (op[10] + c): words[6] + 5 * (op[10] + c) + 5]
                    assert cd[:4] == site
                    named = set(s for k, _b in cases if k is not None for s in k)
                    want = set(key) if key is not None else set(range(n_species)) - named
                    assert set(s for s in range(n_species) if (cd[4] >> s) & 1) == want
                    assert bool((cd[4] >> 31) & 1) == (key is None)
        else:
            assert (ev[6] >= 0) == bool(tail)
            if tail:
                assert info2["routine_ids"]["__otf_tail_" + ir["run_proc"][p][0][1]] == ev[6]
