// kb_otf.cuh -- warp-per-replica step kernel for the otf backend.
//
// The otf base module (kmos/fortran_src/base_otf.f90) re-sums every row of rates_matrix serially on every step
// (update_accum_rate, :687-717) and builds a serial prefix of the chosen row for the site search
// (determine_procsite, :1213-1277): O(number of available events) float64 additions per step whose order is
// part of the result.  A replica therefore cannot split one row over lanes -- but rows are independent:
//   * lane i streams row i of rates_matrix (its own serial chain, the reference's order), records the running
//     sum every KB_OTF_CHUNK entries, and ends with rates_matrix(i, volume+1);
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
#define KB_OTF_RG 4       // rows summed concurrently (one lane each)
#define KB_OTF_STAGE 128  // entries per row and pipeline stage
#define KB_OTF_ROWPAD (KB_OTF_STAGE + 2)  // 16 B-aligned rows whose LDS.128 streams fall into disjoint bank groups
#define KB_OTF_WARP_SMEM (2 * KB_OTF_RG * KB_OTF_ROWPAD * 8)

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
};

template <typename idx_t>
__global__ void __launch_bounds__(128) kb_otf_kernel(const KbOtfParams prm) {
    const int lane = threadIdx.x & 31;
    const int rep = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (rep >= prm.R) return;
    const int P = prm.m.n_proc, C = prm.g.ncells;
    const int nchunk = (C + KB_OTF_CHUNK - 1) / KB_OTF_CHUNK;

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
    r.accum_proc = prm.accum_proc + (size_t)rep * C;  // scratch: [P][nchunk] running sums
    r.lut = prm.lut + (size_t)rep * (prm.m.lut_total > 0 ? prm.m.lut_total : 1);
    const KbScalars s0 = prm.sc[rep];
    r.kmc_time = s0.kmc_time; r.kmc_time_step = s0.kmc_time_step; r.kmc_step = s0.kmc_step;
    r.seed = s0.seed; r.replica = s0.replica; r.status = s0.status;
    for (int i = 0; i < 5; ++i) r.err[i] = s0.err[i];
    KbInterp<idx_t> it(prm.m, prm.g, r);
    extern __shared__ __align__(16) unsigned char kb_otf_smem[];
    double* stage_buf = reinterpret_cast<double*>(kb_otf_smem + (size_t)(threadIdx.x >> 5) * KB_OTF_WARP_SMEM);

    for (long long step = 0; step < prm.nsteps; ++step) {
        int status = __shfl_sync(KB_FULL, r.status, 0);
        if (status != KB_OK) break;
        // -- update_accum_rate: rows in groups of KB_OTF_RG; the whole warp streams the group's next stage into
        //    shared memory with cp.async (coalesced) while lane r < RG adds up the current stage of row g*RG + r
        //    serially, in the reference's order
        for (int g0 = 0; g0 < P; g0 += KB_OTF_RG) {
            const int myrow = g0 + lane;                      // meaningful for lane < RG
            const bool summer = lane < KB_OTF_RG && myrow < P;
            const int my_n = summer ? r.nsites[myrow] : 0;
            int max_n = my_n;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) max_n = max(max_n, __shfl_xor_sync(KB_FULL, max_n, o));
            const int n_stages = (max_n + KB_OTF_STAGE - 1) / KB_OTF_STAGE;
            double tot = 0.0;
            double* marks = r.accum_proc + (size_t)(summer ? myrow : 0) * nchunk;
            const double* rowp[KB_OTF_RG];  // warp-uniform row bases and lengths (only entries < nsites are read)
            int rown[KB_OTF_RG];
#pragma unroll
            for (int rr = 0; rr < KB_OTF_RG; ++rr) {
                rowp[rr] = r.rates_matrix + (size_t)(g0 + rr < P ? g0 + rr : 0) * (C + 1);
                rown[rr] = __shfl_sync(KB_FULL, my_n, rr);
            }
            auto issue = [&](int stage) {
                double* buf = stage_buf + (size_t)(stage & 1) * KB_OTF_RG * KB_OTF_ROWPAD + lane;
                const int base = stage * KB_OTF_STAGE + lane;
#pragma unroll
                for (int rr = 0; rr < KB_OTF_RG; ++rr) {
#pragma unroll
                    for (int k = 0; k < KB_OTF_STAGE; k += 32) {
                        const bool in = base + k < rown[rr];
                        kb_cp_async8(buf + rr * KB_OTF_ROWPAD + k, in ? rowp[rr] + base + k : r.rates_matrix, in);
                    }
                }
                kb_cp_async_commit();
            };
            if (n_stages > 0) issue(0);
            for (int st = 0; st < n_stages; ++st) {
                if (st + 1 < n_stages) { issue(st + 1); kb_cp_async_wait<1>(); } else { kb_cp_async_wait<0>(); }
                __syncwarp();
                if (summer) {
                    const double* buf = stage_buf + (size_t)(st & 1) * KB_OTF_RG * KB_OTF_ROWPAD + lane * KB_OTF_ROWPAD;
                    const int base = st * KB_OTF_STAGE;
                    const int lim = min(KB_OTF_STAGE, my_n - base);
                    for (int k0 = 0; k0 < lim; k0 += KB_OTF_CHUNK) {
                        if (k0 + KB_OTF_CHUNK <= lim) {
                            const double2* b2 = reinterpret_cast<const double2*>(buf + k0);
#pragma unroll
                            for (int k = 0; k < KB_OTF_CHUNK / 2; ++k) {
                                const double2 v = b2[k];
                                tot = __dadd_rn(__dadd_rn(tot, v.x), v.y);
                            }
                            marks[(base + k0) / KB_OTF_CHUNK] = tot;  // accum_rates_proc(base + k0 + CHUNK)
                        } else {
                            for (int k = k0; k < lim; ++k) tot = __dadd_rn(tot, buf[k]);
                        }
                    }
                }
                __syncwarp();  // the buffer is refilled two stages later
            }
            if (summer) r.rates_matrix[(size_t)myrow * (C + 1) + C] = tot;
        }
        __syncwarp();
        if (lane == 0) {
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
    if (lane == 0) {
        KbScalars s = s0;
        s.kmc_time = r.kmc_time; s.kmc_time_step = r.kmc_time_step; s.kmc_step = r.kmc_step; s.status = r.status;
        for (int i = 0; i < 5; ++i) s.err[i] = r.err[i];
        prm.sc[rep] = s;
    }
}
