// kb_smem.cuh -- shared-memory kMC step kernel for sm_100a: one warp steps one replica, the replica's
// lattice and avail-site tables live in shared memory for the whole launch (TMA bulk-staged from/to HBM),
// the warps of a CTA share one copy of the model's per-event lane tables.
//
// Reference loop restated per step (kmos/fortran_src/proclist_generic_subroutines.mpy:22-42):
//   random_number x3            -> per-replica Philox4x32-10 counter stream (kb_common.h)
//   update_accum_rate           base.mpy:603-623   lane q holds nr_of_sites(q)*rates(q); the serial
//                                                  left-to-right float64 recurrence is kept (bit parity)
//   update_clocks               base.mpy:1123-1161
//   update_integ_rate           base.mpy:626-645   lane-parallel over processes
//   determine_procsite          base.mpy:1075-1120 + interval_search_real :1234-1338 as a warp ballot
//   run_proc_nr                 generated put_/take_ routines (kmos/io/__init__.py:305-465, 2219-2409)
//                               flattened by kmos_b200/devtables.py into rounds of per-process list ops;
//                               add_proc/del_proc (base.mpy:211-302) run one op per lane
#pragma once
#include <cuda_runtime.h>

#include "kb_common.h"

struct KbScalars {
    double kmc_time, kmc_time_step;
    int64_t kmc_step;
    uint64_t seed;
    uint32_t replica;
    int32_t status;
    int32_t err[5];
    int32_t pad;
};

struct KbSmemParams {
    // model / geometry
    const int32_t* dev;  // SEC_DEVICE in global memory
    int dev_words;
    int n_proc, spuck, dim;
    int size[3];
    int ncells, volume;
    uint32_t magic_x, magic_xy;  // umulhi(c, magic) == c / Lx  (resp. c / (Lx*Ly)) for all c < ncells
    // batch arrays (global)
    uint8_t* lattice;   // [R][lat_stride]
    int32_t* nsites;    // [R][P]
    uint16_t* p1;       // [R][row_stride]   row_stride = align16(P*ncells*2)/2
    uint16_t* p2;
    const double* rates;
    double* integ;
    int64_t* procstat;
    KbScalars* sc;
    int R;
    long long nsteps;
    // shared-memory layout (bytes)
    int tab_bytes, rep_bytes, off_p2, off_lat, off_ns, off_mbar;
    int lat_stride;       // bytes, multiple of 16
    int plane_bytes;      // bytes of one avail plane, multiple of 16
    int use_bulk;
};

#define KB_FULL 0xffffffffu

__device__ __forceinline__ uint32_t kb_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void kb_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kb_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void kb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kb_smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void kb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KB_DONE_%=;\n"
        "bra KB_WAIT_%=;\n"
        "KB_DONE_%=:\n"
        "}\n" ::"r"(kb_smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void kb_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     kb_smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(kb_smem_addr(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global
__device__ __forceinline__ void kb_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(kb_smem_addr(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void kb_bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void kb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int kb_unpack_s8(uint32_t w, int shift) { return (int)(int8_t)((w >> shift) & 255u); }

struct KbCellCtx {
    int Lx, Ly, Lz, dim, spuck;
    uint32_t magic_x, magic_xy;
    int x, y, z;
    __device__ __forceinline__ void decode(int cell) {
        if (dim >= 3) {
            int LxLy = Lx * Ly;
            z = (int)__umulhi((uint32_t)cell, magic_xy);
            cell -= z * LxLy;
        } else {
            z = 0;
        }
        if (dim >= 2) {
            y = (int)__umulhi((uint32_t)cell, magic_x);
            x = cell - y * Lx;
        } else {
            y = 0;
            x = cell;
        }
    }
    // cell index of (x+dx, y+dy, z+dz) with periodic wrap; |d*| <= L* is checked at batch creation
    __device__ __forceinline__ int cell_at(uint32_t packed) const {
        int xx = x + kb_unpack_s8(packed, 0);
        xx += (xx < 0) ? Lx : 0;
        xx -= (xx >= Lx) ? Lx : 0;
        int c = xx;
        if (dim >= 2) {
            int yy = y + kb_unpack_s8(packed, 8);
            yy += (yy < 0) ? Ly : 0;
            yy -= (yy >= Ly) ? Ly : 0;
            c += Lx * yy;
        }
        if (dim >= 3) {
            int zz = z + kb_unpack_s8(packed, 16);
            zz += (zz < 0) ? Lz : 0;
            zz -= (zz >= Lz) ? Lz : 0;
            c += Lx * Ly * zz;
        }
        return c;
    }
};

// PPL: processes per lane (1: P <= 32, 2: P <= 64)
template <int PPL>
__global__ void kb_smem_kernel(const KbSmemParams prm) {
    extern __shared__ __align__(128) unsigned char kb_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    int32_t* tab = reinterpret_cast<int32_t*>(kb_smem);
    for (int i = threadIdx.x; i < prm.dev_words; i += blockDim.x) tab[i] = prm.dev[i];
    __syncthreads();
    const int rep = blockIdx.x * wpc + warp;
    if (rep >= prm.R) return;  // no block-wide barrier below this line

    const int32_t* events = tab + tab[3];
    const uint32_t* ops = reinterpret_cast<const uint32_t*>(tab + tab[4]);
    const uint32_t* anchors = reinterpret_cast<const uint32_t*>(tab + tab[6]);
    const uint32_t* conds = reinterpret_cast<const uint32_t*>(tab + tab[8]);

    unsigned char* base = kb_smem + prm.tab_bytes + (size_t)warp * prm.rep_bytes;
    uint16_t* p1 = reinterpret_cast<uint16_t*>(base);
    uint16_t* p2 = reinterpret_cast<uint16_t*>(base + prm.off_p2);
    uint8_t* lat = base + prm.off_lat;
    int32_t* nS = reinterpret_cast<int32_t*>(base + prm.off_ns);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(base + prm.off_mbar);

    const int P = prm.n_proc, C = prm.ncells;
    const size_t row_elems = (size_t)prm.plane_bytes / 2;
    uint16_t* g_p1 = prm.p1 + (size_t)rep * row_elems;
    uint16_t* g_p2 = prm.p2 + (size_t)rep * row_elems;
    uint8_t* g_lat = prm.lattice + (size_t)rep * prm.lat_stride;
    int32_t* g_ns = prm.nsites + (size_t)rep * P;

    // ---- stage the replica into shared memory ------------------------------------------------------
    if (prm.use_bulk) {
        if (lane == 0) {
            kb_mbar_init(mbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) {
            kb_mbar_expect_tx(mbar, (uint32_t)(2 * prm.plane_bytes + prm.lat_stride));
            kb_bulk_g2s(p1, g_p1, (uint32_t)prm.plane_bytes, mbar);
            kb_bulk_g2s(p2, g_p2, (uint32_t)prm.plane_bytes, mbar);
            kb_bulk_g2s(lat, g_lat, (uint32_t)prm.lat_stride, mbar);
        }
        kb_mbar_wait(mbar, 0);
    } else {
        const uint4* s1 = reinterpret_cast<const uint4*>(g_p1);
        const uint4* s2 = reinterpret_cast<const uint4*>(g_p2);
        uint4* d1 = reinterpret_cast<uint4*>(p1);
        uint4* d2 = reinterpret_cast<uint4*>(p2);
        for (int i = lane; i < prm.plane_bytes / 16; i += 32) { d1[i] = s1[i]; d2[i] = s2[i]; }
        const uint4* sl = reinterpret_cast<const uint4*>(g_lat);
        uint4* dl = reinterpret_cast<uint4*>(lat);
        for (int i = lane; i < prm.lat_stride / 16; i += 32) dl[i] = sl[i];
    }
    for (int i = lane; i < P; i += 32) nS[i] = g_ns[i];
    __syncwarp();

    // ---- per-lane process registers -----------------------------------------------------------------
    const int q0 = lane, q1 = lane + 32;
    const bool has0 = q0 < P, has1 = (PPL == 2) && (q1 < P);
    const double rate0 = has0 ? prm.rates[(size_t)rep * P + q0] : 0.0;
    const double rate1 = has1 ? prm.rates[(size_t)rep * P + q1] : 0.0;
    double integ0 = has0 ? prm.integ[(size_t)rep * P + q0] : 0.0;
    double integ1 = has1 ? prm.integ[(size_t)rep * P + q1] : 0.0;
    long long ps0 = has0 ? prm.procstat[(size_t)rep * P + q0] : 0;
    long long ps1 = has1 ? prm.procstat[(size_t)rep * P + q1] : 0;

    KbScalars sc = prm.sc[rep];
    double kmc_time = sc.kmc_time, kmc_dt = sc.kmc_time_step;
    long long kmc_step = sc.kmc_step;
    int status = sc.status;
    int err0 = 0, err1 = 0, err2 = 0, err3 = 0, err4 = 0;

    KbCellCtx cc;
    cc.Lx = prm.size[0]; cc.Ly = prm.size[1]; cc.Lz = prm.size[2]; cc.dim = prm.dim; cc.spuck = prm.spuck;
    cc.magic_x = prm.magic_x; cc.magic_xy = prm.magic_xy;
    const uint32_t k0 = (uint32_t)sc.seed, k1 = (uint32_t)(sc.seed >> 32);

    for (long long it = 0; it < prm.nsteps && status == KB_OK; ++it) {
        // -- the step's three uniforms: lane parity picks the Philox slot, results are broadcast
        uint32_t rnd[4];
        kb_philox4x32_10((uint32_t)kmc_step, (uint32_t)((unsigned long long)kmc_step >> 32), sc.replica,
                         (uint32_t)(lane & 1), k0, k1, rnd);
        const uint32_t a0w = __shfl_sync(KB_FULL, rnd[0], 0), a1w = __shfl_sync(KB_FULL, rnd[1], 0);
        const uint32_t a2w = __shfl_sync(KB_FULL, rnd[2], 0), a3w = __shfl_sync(KB_FULL, rnd[3], 0);
        const uint32_t b0w = __shfl_sync(KB_FULL, rnd[0], 1), b1w = __shfl_sync(KB_FULL, rnd[1], 1);
        const double ran_time = (double)(((((uint64_t)a1w << 32) | a0w) >> 11) + 1) * 0x1.0p-53;
        const double ran_proc = (double)((((uint64_t)a3w << 32) | a2w) >> 11) * 0x1.0p-53;
        const double ran_site = (double)((((uint64_t)b1w << 32) | b0w) >> 11) * 0x1.0p-53;

        // -- update_accum_rate: serial float64 recurrence over processes, values exchanged by shuffle
        const int n0 = has0 ? nS[q0] : 0;
        const int n1 = has1 ? nS[q1] : 0;
        const double pr0 = __dmul_rn((double)n0, rate0);
        const double pr1 = __dmul_rn((double)n1, rate1);
        double acc = 0.0, acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
        for (int i = 0; i < P; ++i) {
            const double x = __shfl_sync(KB_FULL, (PPL == 2 && i >= 32) ? pr1 : pr0, i & 31);
            acc = __dadd_rn(acc, x);
            if ((i & 31) == lane) {
                if (PPL == 2 && i >= 32) acc1 = acc; else acc0 = acc;
            }
        }
        const double total = acc;
        if (!(total > 0.0)) { status = KB_DEADLOCK; break; }

        // -- update_clocks / update_integ_rate
        kmc_dt = -log(ran_time) / total;
        kmc_time = __dadd_rn(kmc_time, kmc_dt);
        kmc_step += 1;
        integ0 = __dadd_rn(integ0, __dmul_rn(pr0, kmc_dt));
        if (PPL == 2) integ1 = __dadd_rn(integ1, __dmul_rn(pr1, kmc_dt));

        // -- determine_procsite: first process whose accumulated rate exceeds ran_proc*total
        const double value = __dmul_rn(ran_proc, total);
        const unsigned le0 = __ballot_sync(KB_FULL, has0 && !(value < acc0));
        const unsigned le1 = (PPL == 2) ? __ballot_sync(KB_FULL, has1 && !(value < acc1)) : 0u;
        int pidx = __popc(le0) + __popc(le1);
        if (pidx >= P) {
            // value >= accum(P): the reference's search ends on the last entry and then walks left over
            // entries that are >= their right neighbour (base.mpy:1316-1326)
            const unsigned ge0 = __ballot_sync(KB_FULL, has0 && acc0 >= total);
            const unsigned ge1 = (PPL == 2) ? __ballot_sync(KB_FULL, has1 && acc1 >= total) : 0u;
            pidx = P - (__popc(ge0) + __popc(ge1));
        }
        const int nsel = __shfl_sync(KB_FULL, (PPL == 2 && pidx >= 32) ? n1 : n0, pidx & 31);
        if (nsel <= 0) { status = KB_DEADLOCK; break; }
        int k = (int)__dadd_rn(1.0, __dmul_rn(ran_site, (double)nsel));
        k = min(k, nsel);
        const int cell = (int)p1[pidx * C + k - 1];
        if ((pidx & 31) == lane) {
            if (PPL == 2 && pidx >= 32) ps1 += 1; else ps0 += 1;
        }

        // -- run_proc_nr(pidx+1, site): lattice writes, then rounds of per-process list operations
        const int32_t* ev = events + pidx * KB_DEV_EVENT_STRIDE;
        const int ops_start = ev[0], n_rounds = ev[1], n_writes = ev[2];
        cc.decode(cell);
        if (lane < n_writes) {
            const uint32_t ws = (uint32_t)ev[4 + KB_DEV_MAX_ROUNDS + 2 * lane];
            const uint32_t on = (uint32_t)ev[4 + KB_DEV_MAX_ROUNDS + 2 * lane + 1];
            const int idx = cc.cell_at(ws) * cc.spuck + (int)(ws >> 24) - 1;
            const int found = lat[idx];
            if (found != (int)(on & 255u)) {  // replace_species consistency check (base.mpy:1205)
                status = KB_SPECIES_MISMATCH;
                err0 = (int)(on & 255u); err1 = (int)(on >> 8); err2 = found; err3 = idx + 1; err4 = (int)(kmc_step - 1);
            } else {
                lat[idx] = (uint8_t)(on >> 8);
            }
        }
        int start = 0;
        for (int r = 0; r < n_rounds; ++r) {
            const int endr = ev[4 + r];
            const int i = start + lane;
            if (i < endr) {
                const uint32_t w0 = ops[2 * (ops_start + i)], w1 = ops[2 * (ops_start + i) + 1];
                const int kind = (int)(w0 & 15u), q = (int)((w0 >> 4) & 0xFFFu) - 1;
                const int ncond = (int)(w0 >> 24);
                const int ca = cc.cell_at(anchors[(w0 >> 16) & 255u]);
                const int row = q * C;
                if (kind == KB_KIND_ADD) {
                    bool ok = true;
                    for (int j = 0; j < ncond; ++j) {
                        const uint32_t ci = (w1 >> (8 * j)) & 255u;
                        const uint32_t cs = conds[2 * ci], mask = conds[2 * ci + 1];
                        const int sidx = cc.cell_at(cs) * cc.spuck + (int)(cs >> 24) - 1;
                        const uint32_t sp = lat[sidx];
                        ok = ok && (sp < 32u) && ((mask >> sp) & 1u);
                    }
                    if (ok) {  // add_proc (base.mpy:268-302)
                        const int nq = nS[q];
                        if (nq >= C || p2[row + ca] != 0) {
                            status = KB_CAPACITY;
                        } else {
                            p1[row + nq] = (uint16_t)ca;
                            p2[row + ca] = (uint16_t)(nq + 1);
                            nS[q] = nq + 1;
                        }
                    }
                } else {  // guarded del_proc (base.mpy:211-265)
                    const int pos = p2[row + ca];
                    if (pos != 0) {
                        const int nq = nS[q];
                        const uint16_t last = p1[row + nq - 1];
                        if (pos < nq) {
                            p1[row + pos - 1] = last;
                            p2[row + last] = (uint16_t)pos;
                        }
                        p2[row + ca] = 0;
                        nS[q] = nq - 1;
                    }
                }
            }
            __syncwarp();
            start = endr;
        }
        __syncwarp();  // lattice writes of an event without ops must be visible to the next step
        // a lane-local failure (species mismatch / capacity) stops the replica for every lane
        status = __reduce_max_sync(KB_FULL, status);
    }
    status = __reduce_max_sync(KB_FULL, status);

    // ---- write back -----------------------------------------------------------------------------------
    __syncwarp();
    if (prm.use_bulk) {
        kb_fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            kb_bulk_s2g(g_p1, p1, (uint32_t)prm.plane_bytes);
            kb_bulk_s2g(g_p2, p2, (uint32_t)prm.plane_bytes);
            kb_bulk_s2g(g_lat, lat, (uint32_t)prm.lat_stride);
            kb_bulk_commit_wait();
        }
    } else {
        uint4* s1 = reinterpret_cast<uint4*>(g_p1);
        uint4* s2 = reinterpret_cast<uint4*>(g_p2);
        const uint4* d1 = reinterpret_cast<const uint4*>(p1);
        const uint4* d2 = reinterpret_cast<const uint4*>(p2);
        for (int i = lane; i < prm.plane_bytes / 16; i += 32) { s1[i] = d1[i]; s2[i] = d2[i]; }
        uint4* sl = reinterpret_cast<uint4*>(g_lat);
        const uint4* dl = reinterpret_cast<const uint4*>(lat);
        for (int i = lane; i < prm.lat_stride / 16; i += 32) sl[i] = dl[i];
    }
    for (int i = lane; i < P; i += 32) g_ns[i] = nS[i];
    if (has0) { prm.integ[(size_t)rep * P + q0] = integ0; prm.procstat[(size_t)rep * P + q0] = ps0; }
    if (has1) { prm.integ[(size_t)rep * P + q1] = integ1; prm.procstat[(size_t)rep * P + q1] = ps1; }
    // the failing lane (if any) owns the error tuple
    const unsigned bad = __ballot_sync(KB_FULL, err3 != 0);
    if (bad) {
        const int src = __ffs(bad) - 1;
        err0 = __shfl_sync(KB_FULL, err0, src); err1 = __shfl_sync(KB_FULL, err1, src);
        err2 = __shfl_sync(KB_FULL, err2, src); err3 = __shfl_sync(KB_FULL, err3, src);
        err4 = __shfl_sync(KB_FULL, err4, src);
    }
    if (lane == 0) {
        KbScalars out = sc;
        out.kmc_time = kmc_time; out.kmc_time_step = kmc_dt; out.kmc_step = kmc_step; out.status = status;
        if (bad) { out.err[0] = err0; out.err[1] = err1; out.err[2] = err2; out.err[3] = err3; out.err[4] = err4; }
        prm.sc[rep] = out;
    }
}
