// kb_otf_fast.cuh -- otf production kernel: sub-linear event selection (sm_100a).
//
// The reference's otf base module adds up every live entry of rates_matrix on every step and samples the site
// from a serial prefix of the chosen row (base_otf.f90:687-717, 1213-1277): O(N_sites) per step, and its own
// documentation names that as the backend's known issue, to be replaced by an O(log N) scheme
// (doc/source/topic_guides/otf_backend.rst:198-202).  kb_otf.cuh keeps the O(N) sums because their order is part
// of the bit-exact result and runs them at the HBM roofline; this kernel is the production path:
//
//   * every row of rates_matrix carries block sums over KB_OTFF_BLOCK = 256 consecutive positions (stored in
//     the replica's accum_rates_proc scratch array) next to its row total; add_proc / del_proc /
//     update_rates_matrix adjust the one or two block sums they touch (KbInterp, r.blk != nullptr);
//   * a step reads the P row totals, finds the process, then finds the block by a warp-parallel prefix over
//     the row's <= ncells/256 block sums and the position by a warp-parallel prefix over that block's 256
//     entries: two dependent, coalesced reads of <= 2 KB each instead of the whole row (256x256 lattice: ~5 KB
//     per step instead of 1 MB);
//   * the prefix order is the reference's (position order within a row), so in exact arithmetic the selected
//     (process, site) is the reference's for the same random numbers; in floating point the partial sums are
//     associated differently, i.e. the choice differs only when a random number falls within rounding distance
//     of an interval boundary.  It is therefore NOT the bit-exact path: kmc_time agrees to ~1e-13 per step and
//     trajectories coincide with the exact kernel's until the first such coincidence (tests: identical over
//     thousands of steps, statistically equivalent beyond);
//   * rounding drift of the incrementally maintained sums is bounded by re-adding every block sum and row total
//     from the entries every KB_OTFF_REBUILD steps (and at the start of every launch): what
//     base.reaccumulate_rates_matrix (base_otf.f90:366-387) does on request.
//
// One warp steps one replica; the event itself (guarded dels, update_rates_matrix with gr_<proc> look-ups,
// if-tree adds) is executed by lane 0 through the byte-code interpreter, as in kb_otf.cuh.  Selected with
// kmos_b200_select_kernel(KMOS_B200_KERNEL_OTF_FAST); never chosen automatically.
#pragma once
#include "kb_otf.cuh"
#include "kb_otf_event.cuh"

#define KB_OTFF_SHIFT 8
#define KB_OTFF_BLOCK (1 << KB_OTFF_SHIFT)
#define KB_OTFF_REBUILD 2048
#define KB_OTFF_WARPS 8

// first index k in a[0, m) whose inclusive prefix sum exceeds `value` (lane l owns the contiguous chunk
// [l*per, (l+1)*per), per <= 8), -1 if none; *before = prefix sum in front of k.  Warp-collective.
__device__ __forceinline__ int kb_otff_search(const double* a, int m, double value, double* before) {
    const int lane = threadIdx.x & 31;
    const int per = (m + 31) >> 5;
    const int lo = min(lane * per, m), hi = min(lo + per, m);
    double v[8];
    double local = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = (lo + i < hi) ? a[lo + i] : 0.0;
        local += v[i];
    }
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(KB_FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned hit = __ballot_sync(KB_FULL, value < incl && hi > lo);
    if (!hit) return -1;
    const int src = __ffs(hit) - 1;
    int k = -1;
    double bef = 0.0;
    if (lane == src) {
        double s = incl - local, lastv = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (k < 0 && lo + i < hi) {
                if (value < s + v[i]) { k = lo + i; bef = s; }
                s += v[i];
                lastv = v[i];
            }
        }
        if (k < 0) { k = hi - 1; bef = s - lastv; }  // rounding: the lane's total said "here"
    }
    k = __shfl_sync(KB_FULL, k, src);
    *before = __shfl_sync(KB_FULL, bef, src);
    return k;
}

template <typename idx_t>
__global__ void __launch_bounds__(32 * KB_OTFF_WARPS) kb_otf_fast_kernel(const KbOtfParams prm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rep = blockIdx.x * KB_OTFF_WARPS + warp;
    if (rep >= prm.R) return;  // warps are independent workers: no block-wide barrier below
    const int P = prm.m.n_proc, C = prm.g.ncells;
    const int nblk = (C + KB_OTFF_BLOCK - 1) >> KB_OTFF_SHIFT;
    const size_t lut_stride = prm.m.lut_total > 0 ? prm.m.lut_total : 1;

    KbReplica<idx_t> r;
    r.lattice = prm.lattice + (size_t)rep * prm.lat_stride;
    r.nsites = prm.nsites + (size_t)rep * P;
    r.p1 = reinterpret_cast<idx_t*>(prm.p1) + (size_t)rep * prm.plane_elems;
    r.p2 = reinterpret_cast<idx_t*>(prm.p2) + (size_t)rep * prm.plane_elems;
    r.rates = prm.rates + (size_t)rep * P;
    r.integ = prm.integ + (size_t)rep * P;
    r.accum = prm.accum + (size_t)rep * P;
    r.procstat = prm.procstat + (size_t)rep * P;
    r.rates_matrix = prm.rates_matrix + (size_t)rep * P * (C + 1);
    r.accum_proc = prm.accum_proc + (size_t)rep * C;
    r.blk = r.accum_proc;  // [P][nblk] block sums (P * nblk <= C is checked by the host)
    r.blk_n = nblk;
    r.lut = prm.lut + (size_t)rep * lut_stride;
    const KbScalars s0 = prm.sc[rep];
    r.kmc_time = s0.kmc_time; r.kmc_time_step = s0.kmc_time_step; r.kmc_step = s0.kmc_step;
    r.seed = s0.seed; r.replica = s0.replica; r.status = s0.status;
    for (int i = 0; i < 5; ++i) r.err[i] = s0.err[i];
    KbInterp<idx_t> it(prm.m, prm.g, r);
    __shared__ double s_accum[KB_OTFF_WARPS][64];  // accum_rates of each warp's replica (P <= 64)

    for (long long step = 0; step < prm.nsteps; ++step) {
        if (__shfl_sync(KB_FULL, r.status, 0) != KB_OK) break;
        if ((step & (KB_OTFF_REBUILD - 1)) == 0) {
            // re-add every block sum and row total from the entries (coalesced, the whole warp per block)
            for (int p = 0; p < P; ++p) {
                const double* rm = r.rates_matrix + (size_t)p * (C + 1);
                const int n = r.nsites[p];
                double tot = 0.0;
                for (int b = 0; b < nblk; ++b) {
                    const int base = b << KB_OTFF_SHIFT;
                    double s = 0.0;
                    if (base < n) {
#pragma unroll
                        for (int i = 0; i < KB_OTFF_BLOCK / 32; ++i) {
                            const int k = base + i * 32 + lane;
                            s += (k < n) ? rm[k] : 0.0;
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(KB_FULL, s, o);
                    }
                    if (lane == 0) r.blk[(size_t)p * nblk + b] = s;
                    tot += s;
                }
                if (lane == 0) r.rates_matrix[(size_t)p * (C + 1) + C] = tot;
            }
            __syncwarp();
        }
        // -- update_accum_rate over the maintained row totals, update_clocks, update_integ_rate, process search.
        // Lane l holds the rows l and l + 32: one round trip for all totals, counters and integrals instead of
        // one per process; the accumulation itself is the reference's serial recurrence, repeated by every lane
        // on shuffled values (same operations in the same order: same bits), the search reads it from shared memory.
        const bool h0 = lane < P, h1 = lane + 32 < P;
        const double tot0 = h0 ? r.rates_matrix[(size_t)lane * (C + 1) + C] : 0.0;
        const double tot1 = h1 ? r.rates_matrix[(size_t)(lane + 32) * (C + 1) + C] : 0.0;
        const int ns0 = h0 ? r.nsites[lane] : 0, ns1 = h1 ? r.nsites[lane + 32] : 0;
        const double in0 = h0 ? r.integ[lane] : 0.0, in1 = h1 ? r.integ[lane + 32] : 0.0;
        double* sacc = s_accum[warp];
        double acc = 0.0;
        for (int i = 0; i < P; ++i) {
            const double t = __shfl_sync(KB_FULL, i < 32 ? tot0 : tot1, i & 31);
            acc = (i == 0) ? t : acc + t;
            if ((i & 31) == lane) sacc[i] = acc;
        }
        __syncwarp();
        const double total = acc;  // accum_rates(nr_of_proc)
        if (h0) r.accum[lane] = sacc[lane];
        if (h1) r.accum[lane + 32] = sacc[lane + 32];
        int p = 0;
        double ran_site = 0.0;
        bool alive = total > 0.;
        if (lane == 0) {
            if (!alive) {
                it.fail(KB_DEADLOCK);
            } else {
                double ran_time, ran_proc;
                kb_philox_step(r.seed, r.replica, (uint64_t)r.kmc_step, &ran_time, &ran_proc, &ran_site);
                r.kmc_time_step = -log(ran_time) / total;
                r.kmc_time = r.kmc_time + r.kmc_time_step;
                r.kmc_step = r.kmc_step + 1;
                p = KbInterp<idx_t>::interval_search_real(sacc, P, ran_proc * total);
            }
        }
        if (!alive) continue;  // stopped: the status check at the top of the loop ends the launch for this replica
        const double dt = __shfl_sync(KB_FULL, r.kmc_time_step, 0);
        r.kmc_step = __shfl_sync(KB_FULL, (long long)r.kmc_step, 0);  // every lane may have to report an error
        if (h0) r.integ[lane] = KB_ADD(in0, KB_MUL(tot0, dt));
        if (h1) r.integ[lane + 32] = KB_ADD(in1, KB_MUL(tot1, dt));
        p = __shfl_sync(KB_FULL, p, 0);
        const int psel = p > 0 ? p - 1 : 0;
        const int n = __shfl_sync(KB_FULL, psel < 32 ? ns0 : ns1, psel & 31);
        const double tsel = __shfl_sync(KB_FULL, psel < 32 ? tot0 : tot1, psel & 31);
        if (p == 0 || n <= 0) {
            if (lane == 0) it.fail(KB_DEADLOCK);
            continue;
        }
        const double value = __shfl_sync(KB_FULL, ran_site * tsel, 0);
        // -- determine_procsite: block, then position inside the block
        const double* rm = r.rates_matrix + (size_t)(p - 1) * (C + 1);
        const int nb = (n + KB_OTFF_BLOCK - 1) >> KB_OTFF_SHIFT;
        double before = 0.0;
        int b = kb_otff_search(r.blk + (size_t)(p - 1) * nblk, nb, value, &before);
        int k = -1;
        if (b >= 0) {
            const int base = b << KB_OTFF_SHIFT;
            double bef2 = 0.0;
            k = kb_otff_search(rm + base, min(KB_OTFF_BLOCK, n - base), value - before, &bef2);
            if (k >= 0) k += base;
        }
        if (k < 0) {
            // value >= what the sums add up to (rounding of the maintained sums): the reference's search ends on
            // the last entry and walks left over zero rates (base.mpy:1316-1326)
            k = n - 1;
            if (lane == 0) while (k > 0 && !(rm[k] > 0.)) --k;
            k = __shfl_sync(KB_FULL, k, 0);
        }
        if (prm.lanes) {
            const int cell = (int)r.p1[(size_t)(p - 1) * C + k];  // one address for the whole warp
            kb_otff_event(it, r, prm.m, prm.g, prm.lanes, p, cell, lane);
        } else {
            if (lane == 0) {
                const int cell = (int)r.p1[(size_t)(p - 1) * C + k];
                it.run_proc_nr(p, cell);
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        KbScalars s = s0;
        s.kmc_time = r.kmc_time; s.kmc_time_step = r.kmc_time_step; s.kmc_step = r.kmc_step; s.status = r.status;
        for (int i = 0; i < 5; ++i) s.err[i] = r.err[i];
        prm.sc[rep] = s;
    }
}
