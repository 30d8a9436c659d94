// kb_otf_event.cuh -- run_proc_<proc>(cell) of an otf model spread over the lanes of a warp (sm_100a).
//
// Used by the production kernel (kb_otf_fast.cuh).  Linked into the exact kernel (kb_otf.cuh) as well it was
// bit-exact on every otf test but cost that kernel 7 % (register pressure on its summation loops) for an event that
// is 3 % of its step: left out there.  Layout of the lane tables and the argument why the result is the serial
// one: kmos_b200/devtables.py, compile_otf_tables.
#pragma once
#include "kb_interp.h"
#include "kb_smem.cuh"

#define KB_OTFF_OP 12  // words per lane-table op (devtables.py, OTF_OP_WORDS)

// The first error any lane ran into becomes lane 0's (the lane whose status is stored): lanes execute statements
// in textual order, so the lowest lane is the reference's first.
template <typename idx_t>
__device__ __forceinline__ void kb_otff_merge_status(KbReplica<idx_t>& r, int lane) {
    const unsigned bad = __ballot_sync(KB_FULL, r.status != KB_OK);
    if (!bad) return;
    const int src = __ffs(bad) - 1;
    const int st = __shfl_sync(KB_FULL, r.status, src);
    int e[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) e[i] = __shfl_sync(KB_FULL, r.err[i], src);
    if (lane == 0 && src != 0) {
        r.status = st;
#pragma unroll
        for (int i = 0; i < 5; ++i) r.err[i] = e[i];
    }
}

// run_proc_<proc>(cell) with the statements of its first three blocks spread over the lanes (layout and
// rationale: devtables.py, compile_otf_tables).  Statements on one process keep their textual order (rank among
// the lanes that name it); statements on different processes touch different rows of avail_sites and
// rates_matrix.  Bit-identical to lane 0 interpreting the routine.
template <typename idx_t>
__device__ __forceinline__ void kb_otff_event(KbInterp<idx_t>& it, KbReplica<idx_t>& r, const KbModelView& m,
                                              const KbGeom& g, const int32_t* T, int proc, int cell, int lane) {
    const int32_t* ev = T + T[3] + 8 * (proc - 1);
    const int32_t* ops = T + T[4];
    int base[4];
    it.cell_coords(cell, base);
    base[3] = 0;  // the routine is called on the cell: site types are the statements' own fourth offsets
    if (lane == 0) r.procstat[proc - 1]++;
    const unsigned lt_mask = (1u << lane) - 1u;
    // Loads nobody waits for until the end of the event: the back-pointers the update block will test and the
    // lattice rows its gr_<proc> functions count bystanders in are on their way while the dels run (the
    // statements load them again at their proper time -- from L1).  Values: at most 2^17 resp. 255.
    uint32_t warm = 0;
    if (lane < ev[5] && lane < 24) {
        const int32_t* op = ops + KB_OTFF_OP * (ev[4] + lane);
        const typename KbInterp<idx_t>::Site s = it.site_of(base, op + 1);
        warm = (uint32_t)r.p2[(size_t)(op[0] - 1) * g.ncells + s.cell];
    } else if (lane >= 24 && lane < 31) {
        const int j = lane - 24;  // rows y-2 .. y+2 of the lattice, planes z-1, z+1
        const int32_t off[4] = {0, (j < 5 && m.dim >= 2) ? j - 2 : 0, (j >= 5 && m.dim >= 3) ? 2 * j - 11 : 0, 1};
        const typename KbInterp<idx_t>::Site s = it.site_of(base, off);
        warm = r.lattice[s.cell * m.spuck];
    }
    // -- if (can_do(q, site)) del_proc(q, site)
    for (int b0 = 0; b0 < ev[1]; b0 += 32) {
        const bool on = b0 + lane < ev[1];
        const int32_t* op = ops + KB_OTFF_OP * (ev[0] + b0 + (on ? lane : 0));
        const int q = op[0];
        const typename KbInterp<idx_t>::Site s = it.site_of(base, op + 1);
        const bool fire = on && it.pos_of(q, s.cell, s.n) != 0;  // every lane's guard in one DRAM round trip
        const unsigned peers = __match_any_sync(KB_FULL, fire ? q : -1 - lane);
        const int rank = __popc(peers & lt_mask);
        const int maxrank = __reduce_max_sync(KB_FULL, fire ? rank : 0);
        for (int k = 0; k <= maxrank; ++k) {
            // re-checked at its turn: an earlier del of the same process may have been on the same site
            if (fire && rank == k && it.pos_of(q, s.cell, s.n)) it.del_proc(q, s.cell, s.n);
            __syncwarp();
        }
    }
    // -- replace_species(site, old, new)
    for (int b0 = 0; b0 < ev[3]; b0 += 32) {
        if (b0 + lane < ev[3]) {
            const int32_t* op = ops + KB_OTFF_OP * (ev[2] + b0 + lane);
            const typename KbInterp<idx_t>::Site s = it.site_of(base, op + 1);
            it.replace_species(s.cell, s.n, op[0], op[5]);
        }
    }
    __syncwarp();
    // -- if (can_do(q, site)) update_rates_matrix(q, site, gr_q(cell'))
    for (int b0 = 0; b0 < ev[5]; b0 += 32) {
        const bool on = b0 + lane < ev[5];
        const int32_t* op = ops + KB_OTFF_OP * (ev[4] + b0 + (on ? lane : 0));
        const int q = op[0];
        const typename KbInterp<idx_t>::Site s = it.site_of(base, op + 1);
        const int pos = on ? it.pos_of(q, s.cell, s.n) : 0;
        const bool fire = pos != 0;
        double rate = 0.0, old = 0.0;
        double* rm = r.rates_matrix + (size_t)(q - 1) * (g.ncells + 1);
        if (fire) {
            old = rm[pos - 1];
            rate = it.eval_gr(op[5], base, op + 6);
            rm[pos - 1] = rate;
        }
        // block sum and row total of one process: in textual order, like update_rates_matrix one by one
        const unsigned peers = __match_any_sync(KB_FULL, fire ? q : -1 - lane);
        const int rank = __popc(peers & lt_mask);
        const int maxrank = __reduce_max_sync(KB_FULL, fire ? rank : 0);
        for (int k = 0; k <= maxrank; ++k) {
            if (fire && rank == k) {
                if (r.blk) r.blk[(size_t)(q - 1) * r.blk_n + ((pos - 1) >> 8)] += rate - old;
                rm[g.ncells] = KB_SUB(KB_ADD(rm[g.ncells], rate), old);
            }
            __syncwarp();
        }
    }
    if (warm >= 0x40000000u) r.status = KB_BAD_MODEL;  // never: ends the life of the warm-up loads
    // -- the if-tree of add_proc(q, site, gr_q(cell')) statements, flattened: each statement with the case labels
    // on its path as conditions; all conditions and rates at once, the appends of one process in textual order
    const int32_t* conds = T + T[6];
    for (int b0 = 0; b0 < ev[7]; b0 += 32) {
        const bool on = b0 + lane < ev[7];
        const int32_t* op = ops + KB_OTFF_OP * (ev[4] + ev[5] + b0 + (on ? lane : 0));
        const int q = op[0];
        bool fire = on;
        for (int c = 0; c < op[11] && fire; ++c) {
            const int32_t* cd = conds + 5 * (op[10] + c);
            const int sp = it.species_at(base, cd);
            fire = ((uint32_t)cd[4] >> (sp >= 0 ? sp : 31)) & 1u;
        }
        const typename KbInterp<idx_t>::Site s = it.site_of(base, op + 1);
        const double rate = fire ? it.eval_gr(op[5], base, op + 6) : 0.0;
        const unsigned peers = __match_any_sync(KB_FULL, fire ? q : -1 - lane);
        const int rank = __popc(peers & lt_mask);
        const int maxrank = __reduce_max_sync(KB_FULL, fire ? rank : 0);
        for (int k = 0; k <= maxrank; ++k) {
            if (fire && rank == k) it.add_proc(q, s.cell, s.n, rate);
            __syncwarp();
        }
    }
    kb_otff_merge_status(r, lane);
    // -- add_proc(q, site, gr_q(cell')) and the select case nests around them: byte-code, lane 0
    if (lane == 0 && ev[6] >= 0 && r.status != KB_BAD_MODEL) it.exec(ev[6], base);
    __syncwarp();
}
