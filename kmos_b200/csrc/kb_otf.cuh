// kb_otf.cuh -- warp-per-replica step kernel for the otf backend.
//
// The otf base module (kmos/fortran_src/base_otf.f90) re-sums every row of rates_matrix serially on every step
// (update_accum_rate, :687-717) and builds a serial prefix of the chosen row for the site search
// (determine_procsite, :1213-1277): O(number of available events) float64 additions per step whose order is
// part of the result.  A replica therefore cannot split one row over lanes -- but rows are independent:
//   * a CTA holds KB_OTF_WARPS replicas; their P rows each are the CTA's "chains".  All warps stream the chains'
//     next stages into a shared-memory ring with cp.async (coalesced), while warp 0 adds them up -- lane c owns
//     chain c: its own serial sum in the reference's order, 32 chains per instruction -- records the running sum
//     every KB_OTF_CHUNK entries, and ends with rates_matrix(i, volume+1);
//   * the process search runs on the P row totals; the site search finds the chunk by bisection over the
//     recorded running sums (they ARE entries of the reference's accum_rates_proc array) and finishes with at
//     most KB_OTF_CHUNK additions -- same comparisons, same index, no second O(n) pass;
//   * the event itself (guarded dels, update_rates_matrix with gr_<proc> look-ups, if-tree adds;
//     kmos/io/__init__.py:3328-3596) is executed by lane 0 through the byte-code interpreter (kb_interp.h).
// State stays in the canonical HBM layout.
#pragma once
#include "kb_interp.h"
#include "kb_smem.cuh"

#define KB_OTF_CHUNK 64
#define KB_OTF_WARPS 8                     // replicas (= warps) per CTA
#define KB_OTF_STAGE KB_OTF_CHUNK          // entries per chain and pipeline stage
#define KB_OTF_NBUF 3                      // stage ring: two stages in flight while one is summed
#define KB_OTF_ROWPAD (KB_OTF_STAGE + 2)   // 16 B-aligned rows; the 32 LDS.128 streams split into 4 conflict-free phases
#define KB_OTF_SMEM (KB_OTF_NBUF * 32 * KB_OTF_ROWPAD * 8)
#define KB_OTF_UNITS ((64 + KB_OTF_WARPS - 2) / (KB_OTF_WARPS - 1))  // (chain, half) units per copying warp

__device__ __forceinline__ void kb_cp_async8(void* dst_smem, const void* src_gmem, bool pred) {
    const int sz = pred ? 8 : 0;  // src-size 0: zero fill, no global access
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(kb_smem_addr(dst_smem)), "l"(src_gmem), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void kb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void kb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct KbOtfParams {
    KbModelView m;
    KbGeom g;
    int R, lat_stride;
    size_t plane_elems;
    uint8_t* lattice;
    void *p1, *p2;
    int32_t* nsites;
    double *rates, *integ, *accum;
    int64_t* procstat;
    KbScalars* sc;
    double *rates_matrix, *accum_proc, *lut;
    long long nsteps;
    const int32_t* lanes = nullptr;  // kb_otf_fast.cuh: the model's otf lane tables (devtables.compile_otf_tables)
};

template <typename idx_t>
__global__ void __launch_bounds__(32 * KB_OTF_WARPS) kb_otf_kernel(const KbOtfParams prm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rep0 = blockIdx.x * KB_OTF_WARPS;
    const int rep = rep0 + warp;
    const bool have = rep < prm.R;  // warps without a replica still help with the loads
    const int P = prm.m.n_proc, C = prm.g.ncells;
    const int nchunk = (C + KB_OTF_CHUNK - 1) / KB_OTF_CHUNK;
    const size_t lut_stride = prm.m.lut_total > 0 ? prm.m.lut_total : 1;

    extern __shared__ __align__(16) unsigned char kb_otf_smem[];
    double* ring = reinterpret_cast<double*>(kb_otf_smem);
    __shared__ const double* ch_ptr[32];
    __shared__ int ch_n[32];
    __shared__ int ch_max;
    __shared__ int rep_status[KB_OTF_WARPS];

    KbReplica<idx_t> r;
    const int rsafe = have ? rep : 0;
    r.lattice = prm.lattice + (size_t)rsafe * prm.lat_stride;
    r.nsites = prm.nsites + (size_t)rsafe * P;
    r.p1 = reinterpret_cast<idx_t*>(prm.p1) + (size_t)rsafe * prm.plane_elems;
    r.p2 = reinterpret_cast<idx_t*>(prm.p2) + (size_t)rsafe * prm.plane_elems;
    r.rates = prm.rates + (size_t)rsafe * P;
    r.integ = prm.integ + (size_t)rsafe * P;
    r.accum = prm.accum + (size_t)rsafe * P;
    r.procstat = prm.procstat + (size_t)rsafe * P;
    r.rates_matrix = prm.rates_matrix + (size_t)rsafe * P * (C + 1);
    r.accum_proc = prm.accum_proc + (size_t)rsafe * C;  // scratch: [P][nchunk] running sums
    r.lut = prm.lut + (size_t)rsafe * lut_stride;
    const KbScalars s0 = prm.sc[rsafe];
    r.kmc_time = s0.kmc_time; r.kmc_time_step = s0.kmc_time_step; r.kmc_step = s0.kmc_step;
    r.seed = s0.seed; r.replica = s0.replica; r.status = have ? s0.status : KB_BAD_MODEL;
    for (int i = 0; i < 5; ++i) r.err[i] = s0.err[i];
    KbInterp<idx_t> it(prm.m, prm.g, r);
    const int n_chains = KB_OTF_WARPS * P;

    for (long long step = 0; step < prm.nsteps; ++step) {
        if (lane == 0) rep_status[warp] = r.status;
        if (__syncthreads_and(__shfl_sync(KB_FULL, r.status, 0) != KB_OK)) break;  // every replica here has stopped
        // -- update_accum_rate over the CTA's chains, 32 at a time
        for (int c0 = 0; c0 < n_chains; c0 += 32) {
            if (warp == 0) {
                const int c = c0 + lane, j = c / P, i = c - j * P;
                const bool live = c < n_chains && rep0 + j < prm.R && rep_status[j] == KB_OK;
                const int n = live ? prm.nsites[(size_t)(rep0 + j) * P + i] : 0;
                ch_ptr[lane] = prm.rates_matrix + ((size_t)(live ? rep0 + j : 0) * P + (live ? i : 0)) * (C + 1);
                ch_n[lane] = n;
                int mx = n;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(KB_FULL, mx, o));
                if (lane == 0) ch_max = mx;
            }
            __syncthreads();
            const int n_stages = (ch_max + KB_OTF_STAGE - 1) / KB_OTF_STAGE;
            // Warp 0 only sums; the other warps only copy.  A stage is 64 (chain, half) units of 32 entries;
            // copier w takes units w-1, w-1+7, ... and keeps their source, length and ring offset in registers.
            const double* u_src[KB_OTF_UNITS];
            int u_n[KB_OTF_UNITS];
            uint32_t u_dst[KB_OTF_UNITS];
#pragma unroll
            for (int i = 0; i < KB_OTF_UNITS; ++i) {
                const int u = warp - 1 + (KB_OTF_WARPS - 1) * i;
                const bool ok = warp > 0 && u < 64;
                const int chain = ok ? u >> 1 : 0, half = u & 1;
                u_src[i] = ch_ptr[chain] + half * 32 + lane;
                u_n[i] = ok ? ch_n[chain] - half * 32 - lane : 0;  // entries of this lane's column still to copy
                u_dst[i] = kb_smem_addr(ring + (size_t)chain * KB_OTF_ROWPAD + half * 32 + lane);
            }
            const int my_n = ch_n[lane];
            double tot = 0.0;
            double* marks = nullptr;
            if (warp == 0) {
                const int c = c0 + lane, j = c / P, i = c - j * P;
                marks = prm.accum_proc + (size_t)(rep0 + j < prm.R ? rep0 + j : 0) * C + (size_t)i * nchunk;
            }
            int ring_w = 0;  // ring slot the next copied stage goes to
            for (int st = -2; st < n_stages; ++st) {
                if (st >= 0) {
                    if (warp != 0) kb_cp_async_wait<1>();  // this thread's part of stage st has landed
                    __syncthreads();                       // ... everybody's has, and warp 0 is done with stage st-1
                }
                if (warp != 0) {
                    const int stage = st + 2;  // goes to the slot of stage st-1
                    if (stage < n_stages) {
                        const uint32_t slot = (uint32_t)ring_w * (32 * KB_OTF_ROWPAD * 8);
                        const int base = stage * KB_OTF_STAGE;
#pragma unroll
                        for (int i = 0; i < KB_OTF_UNITS; ++i) {
                            // only live entries are read; the rest of the slot is never summed
                            if (base < u_n[i])
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(u_dst[i] + slot),
                                             "l"(u_src[i] + base)
                                             : "memory");
                        }
                    }
                    kb_cp_async_commit();  // an empty group keeps the wait counts aligned
                    ring_w = ring_w == KB_OTF_NBUF - 1 ? 0 : ring_w + 1;
                } else if (st >= 0) {
                    const double* buf = ring + ((size_t)(st % KB_OTF_NBUF) * 32 + lane) * KB_OTF_ROWPAD;
                    const int lim = my_n - st * KB_OTF_STAGE;
                    if (lim >= KB_OTF_STAGE) {
                        const double2* b2 = reinterpret_cast<const double2*>(buf);
#pragma unroll
                        for (int k = 0; k < KB_OTF_STAGE / 2; ++k) {
                            const double2 v = b2[k];
                            tot = __dadd_rn(__dadd_rn(tot, v.x), v.y);
                        }
                        marks[st] = tot;  // accum_rates_proc(base + CHUNK)
                    } else {
                        for (int k = 0; k < lim; ++k) tot = __dadd_rn(tot, buf[k]);
                    }
                }
            }
            kb_cp_async_wait<0>();
            if (warp == 0 && c0 + lane < n_chains && my_n >= 0) {
                const int c = c0 + lane, j = c / P;
                if (rep0 + j < prm.R && rep_status[j] == KB_OK) const_cast<double*>(ch_ptr[lane])[C] = tot;
            }
            __syncthreads();  // totals and marks are visible to the replicas' warps; descriptors may be rewritten
        }
        if (lane == 0 && r.status == KB_OK) {
            double acc = 0.0;
            for (int i = 0; i < P; ++i) {
                const double tot = r.rates_matrix[(size_t)i * (C + 1) + C];
                acc = (i == 0) ? tot : __dadd_rn(acc, tot);
                r.accum[i] = acc;
            }
            const double total = r.accum[P - 1];
            if (!(total > 0.)) {
                it.fail(KB_DEADLOCK);
            } else {
                double ran_time, ran_proc, ran_site;
                kb_philox_step(r.seed, r.replica, (uint64_t)r.kmc_step, &ran_time, &ran_proc, &ran_site);
                r.kmc_time_step = -log(ran_time) / total;
                r.kmc_time = __dadd_rn(r.kmc_time, r.kmc_time_step);
                r.kmc_step = r.kmc_step + 1;
                it.update_integ_rate();
                // determine_procsite (base_otf.f90:1213-1277)
                const int p = KbInterp<idx_t>::interval_search_real(r.accum, P, __dmul_rn(ran_proc, total));
                if (p == 0 || r.nsites[p - 1] <= 0) {
                    it.fail(KB_DEADLOCK);
                } else {
                    const int n = r.nsites[p - 1];
                    const double* rm = r.rates_matrix + (size_t)(p - 1) * (C + 1);
                    const double* marks = r.accum_proc + (size_t)(p - 1) * nchunk;
                    const double value = __dmul_rn(ran_site, rm[C]);  // accum_rates_proc(n) == the row total
                    // chunk whose closing running sum is the first one above `value`
                    const int full = n / KB_OTF_CHUNK;
                    int lo = 0, hi = full;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (value < marks[mid]) hi = mid; else lo = mid + 1;
                    }
                    double acc2 = lo > 0 ? marks[lo - 1] : 0.0;
                    int k = lo * KB_OTF_CHUNK;  // entries [0, k) have running sums <= value
                    int found = 0;
                    for (; k < n; ++k) {
                        acc2 = (k == 0) ? rm[0] : __dadd_rn(acc2, rm[k]);
                        if (value < acc2) { found = k + 1; break; }
                    }
                    if (!found) {
                        // value >= accum_rates_proc(n): the reference ends on n and walks left over entries that
                        // are >= their right neighbour, i.e. over trailing zero rates (base.mpy:1316-1326)
                        found = n;
                        while (found > 1 && !(rm[found - 1] > 0.)) --found;
                    }
                    if (!(rm[C] > 0.)) {
                        it.fail(KB_DEADLOCK);
                    } else {
                        const int cell = (int)r.p1[(size_t)(p - 1) * C + found - 1];
                        it.run_proc_nr(p, cell);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0 && have) {
        KbScalars s = s0;
        s.kmc_time = r.kmc_time; s.kmc_time_step = r.kmc_time_step; s.kmc_step = r.kmc_step; s.status = r.status;
        for (int i = 0; i < 5; ++i) s.err[i] = r.err[i];
        prm.sc[rep] = s;
    }
}
