// kb_gen.cuh -- skeleton of the exporter-generated per-model step kernel (sm_100a).
//
// kmos is a code generator: run_proc_nr and the put_/take_ routines it calls exist only as generated,
// model-specific code (kmos/io/__init__.py:305-465 write_proclist_run_proc_nr_smart, :2219-2409
// write_proclist_put_take, :2568-2655 _write_optimal_iftree).  kmos_b200/codegen.py emits the CUDA counterpart,
// proclist_<model>.cu: the model's constants (process count, sites per cell, neighbour offsets, how many
// lanes step one replica), every process' replace_species calls and its guarded del_proc / if-tree add_proc
// calls scheduled into rounds, as static descriptor tables, and the instantiation of this header:
//
//   do_kmc_steps loop        proclist_generic_subroutines.mpy:1-44     kb_gen_kernel<M>
//   update_accum_rate        base.mpy:603-623    packed non-zero products, serial float64 chain per lane
//   update_clocks            base.mpy:1123-1161
//   update_integ_rate        base.mpy:626-645
//   determine_procsite       base.mpy:1075-1120, interval_search_real :1234-1338 as a ballot per lane group
//   add_proc / del_proc      base.mpy:211-302    KbGenCtx::round<>
//   replace_species          base.mpy:1187-1231
//
// Lane groups.  A replica is stepped by a group of LPR lanes (32, 16 or 8: the generator picks the smallest
// group whose rounds are still bounded by the model's dependency chains, not by the group size -- RuO2: 32
// ops per event in 3.6 rounds either way), so one warp steps 32/LPR replicas in lock step and every
// instruction of the step loop is shared by them: different events are different *rows of the tables*, not
// different code, which is why the events are descriptor rows and not unrolled cases (an unrolled RuO2 is
// 196 KB of SASS; with every warp in a different case the instruction cache thrashes: measured IPC 1.3).
//
// State.  Plane 2 of avail_sites (one uint16 per exclusivity class and cell: member << 13 | position, see
// kmos_b200/devtables.py) and the lattice stay in shared memory for the whole work item; plane 1 (one list per
// process) stays in HBM/L2, written through.  What del_proc needs from a list is its *last* element
// (base.mpy:246-254); the group keeps the top four positions of every list in a shared-memory window
// (8 bytes per process, position p in slot p & 3, valid range [lo, nr_of_sites) with lo in the high half of
// the process' nr_of_sites word), so a del reads the element it moves from shared memory and goes to L2 only
// when the window has run empty -- then it fetches the aligned 8-byte chunk holding the last element, which
// refills the window.  add_proc pushes into the window.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "kb_smem.cuh"

#define KB_GEN_ABI 4
#define KB_GEN_MAX_COND 4
#define KB_GEN_KIND_ADD 0x80000000u
#define KB_GEN_ZEROS 16  // leading zeros in front of the packed products (longest straight-line chain)

// ---- static description of a generated module (host side) ----------------------------------------------
struct KbGenOpDesc {  // one list operation, in table order
    uint8_t is_add, q, cls, member, aoff, ncond;
    uint8_t coff[KB_GEN_MAX_COND], cn[KB_GEN_MAX_COND];
    uint16_t cmask[KB_GEN_MAX_COND];
};
struct KbGenRoundDesc {  // one round: `count` (<= LPR) ops from `first_op`
    int32_t first_op, count, nc, kind;
};
struct KbGenEventDesc {  // one process: its rounds
    int32_t first_round, n_rounds;
};
struct KbGenInfo {
    int32_t abi, n_proc, n_species, spuck, dim, n_off, n_classes, n_ops, n_rounds;
    // operand table: [A: one word per op][B: bw probe words per op][round words][event rows][write rows]
    // [one mbarrier per warp][nbT]
    int32_t bw, off_b, off_rd, off_ev, off_wr, ops_bytes, max_threads, lpr;
    uint64_t model_hash;  // FNV-1a of the int32 model blob the code was generated from
    const char* name;
    const KbGenOpDesc* ops;
    const KbGenRoundDesc* rounds;
    const KbGenEventDesc* events;
    const int8_t* offsets;       // [n_off][3]
    const uint32_t* writes;      // [n_proc][4]: off_id | n<<8 | old<<16 | new<<24, 0 = none
    const uint8_t* proc_cls;     // [n_proc] exclusivity class
    const uint8_t* proc_member;  // [n_proc] member tag (1..7)
};

// geometry-dependent layout, computed by kmos_b200_gen_plan
struct KbGenPlan {
    int32_t size[3], ncells, R, device;
    // compact image in HBM: [plane 1: n_proc lists of `cap` uint16][plane 2: n_classes x ncells uint16]
    int32_t cap, off_p2, img_bytes, stage_off, stage_bytes, lat_stride;
    // shared memory: [table][per-warp blocks: 32/lpr replica blocks]; offsets inside a replica block
    int32_t tab_bytes, nbt_off, mbar_off, rep_bytes, warp_bytes, sm_lat, sm_ns, sm_win, sm_prod;
    int32_t lpr, wpc, ctas_per_sm, smem_bytes, regs, sm_count;
    int32_t replicas_per_cta;
};

struct KbGenParams {
    const uint32_t* tab;  // device copy of the table built by kmos_b200_gen_build_tables
    int tab_bytes, nbt_off, mbar_off;
    int ncells, cap;
    uint8_t* lattice;      // [R][lat_stride]
    int32_t* nsites;       // [R][P]
    unsigned char* image;  // [R][img_bytes]
    const double* rates;
    double* integ;
    int64_t* procstat;
    KbScalars* sc;
    const uint32_t* writes;  // device copy of KbGenInfo::writes
    int R;
    long long nsteps;
    int* work_counter;  // [1 + teams]: item counter, then one epoch counter per team of 32/lpr replicas
    int* done;
    int n_items, n_teams;
    long long chunk;
    int rep_bytes, warp_bytes, sm_lat, sm_ns, sm_win, sm_prod;
    int stage_off, stage_bytes, lat_stride, img_bytes;
    int use_bulk;
};

#if defined(__CUDACC__)
// ---- constant-table loads: the operand and neighbour tables never change during a launch, so these asm
// statements carry no memory clobber and the compiler may hoist them over the stores of earlier rounds
__device__ __forceinline__ uint4 kb_ldc128(uint32_t a) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 kb_ldc64(uint32_t a) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t kb_ldc32(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t kb_ldc16(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t kb_shr(uint32_t x, uint32_t s) {  // shift amounts >= 32 give 0 (PTX shr)
    uint32_t v;
    asm("shr.u32 %0, %1, %2;" : "=r"(v) : "r"(x), "r"(s));
    return v;
}

struct KbGenOpB {  // up to 4 if-tree probes: column (byte 0), site (byte 1), species mask (high half)
    uint32_t c[KB_GEN_MAX_COND];
};

// the replica's mutable shared-memory state (class planes, nr_of_sites, windows, lattice) by shared address
__device__ __forceinline__ uint32_t kb_lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t kb_lds16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t kb_lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void kb_sts64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void kb_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void kb_sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void kb_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// del_proc on a list whose window has run empty: fetch the aligned chunk of four positions that holds the
// last element, refill the window with it, return the last element.  Rare, hence out of line.
// (inlined, ptxas predicates it into every round: 12 issue slots per round, measured -6.7 %)
__device__ __noinline__ uint32_t kb_gen_refill(const unsigned char* chunk, const uint32_t wina, const int nq) {
    const uint2 ch = *reinterpret_cast<const uint2*>(chunk);
    kb_sts64(wina, ch);
    const uint32_t w32 = ((nq - 1) & 2) ? ch.y : ch.x;
    return ((nq - 1) & 1) ? (w32 >> 16) : (w32 & 0xffffu);
}

// Everything an event works with.  M: the generated model traits.  All members are per lane; lanes of one
// group hold the same values.
template <class M>
struct KbGenCtx {
    uint32_t wb;         // shared address of the replica's block: class planes at 0
    uint32_t wns;        // ... of its nr_of_sites words (low half: nr_of_sites, high half: window start lo)
    uint32_t win;        // ... of its list windows (8 bytes per process)
    uint32_t lat;        // ... of its lattice copy
    unsigned char* p1;   // the replica's lists in HBM/L2 (list of process q at byte q*cap*2)
    uint32_t tab0;       // shared address of the operand table
    uint32_t tabA;       // ... + sl*4: this lane's operand word of a round whose first op sits at offset 0
    uint32_t tabB;       // ... + OFF_B + sl*4*BW: this lane's probe words
    uint32_t nbT;        // shared address of the neighbour table: row = cell, column = offset, value = 2*cell'
    uint32_t nbrow;      // row of the selected cell
    int sl, C;
    uint32_t capH;       // cap / 2: list byte offset of process q = (4*q) * capH
    uint32_t pl2;        // bytes of one class plane (2 * ncells)
    uint32_t bad;        // bit 0: capacity, bit 1+i: write i found another species

    __device__ __forceinline__ uint16_t* list_at(uint32_t byte_off) const {
        return reinterpret_cast<uint16_t*>(p1 + byte_off);
    }
    __device__ __forceinline__ uint32_t lat_index(uint32_t c2) const {
        return (M::SPUCK % 2 == 0) ? c2 * (M::SPUCK / 2) : (c2 * M::SPUCK) >> 1;
    }
    // operands of the round whose first op has index i (lane sl: op i + sl)
    __device__ __forceinline__ uint32_t ldA(const uint32_t i) const { return kb_ldc32(tabA + 4u * i); }
    __device__ __forceinline__ KbGenOpB ldB(const uint32_t i) const {
        KbGenOpB b;
        b.c[0] = b.c[1] = b.c[2] = b.c[3] = 0;
        if (M::BW == 1) {
            b.c[0] = kb_ldc32(tabB + 4u * i);
        } else if (M::BW == 2) {
            const uint2 v = kb_ldc64(tabB + 8u * i);
            b.c[0] = v.x; b.c[1] = v.y;
        } else if (M::BW == 4) {
            const uint4 v = kb_ldc128(tabB + 16u * i);
            b.c[0] = v.x; b.c[1] = v.y; b.c[2] = v.z; b.c[3] = v.w;
        }
        return b;
    }

    // One round: lane sl < count executes op sl of its group's round.  a = 2 * neighbour column | class << 8 |
    // (4 * process) << 16 | member << 24 | (add ? 1 << 31 : 0): 4 bytes per op keep RuO2's 1496 ops at 6 KB.
    //   guarded del_proc (base.mpy:211-265): registered iff the class entry carries this op's member tag
    //   add_proc (base.mpy:268-302) after the if-tree probes of its leaf (io/__init__.py:2568-2655)
    // Both as one predicated sequence: the groups of a warp (and the lanes of a group) mix them freely.
    __device__ __forceinline__ void round(const int count, const uint32_t a, const KbGenOpB& b) {
        const bool valid = sl < count;
        const uint32_t ca2 = kb_ldc16(nbrow + __byte_perm(a, 0u, 0x4440u));
        bool ok = valid;
#pragma unroll
        for (int j = 0; j < M::BW; ++j) {
            const uint32_t w = b.c[j];  // unused probe slots are always true (see kb_gen_fill_tables)
            const uint32_t cc2 = kb_ldc16(nbrow + __byte_perm(w, 0u, 0x4440u));
            const uint32_t sp = kb_lds8(lat + __byte_perm(w, 0u, 0x4441u) + lat_index(cc2));
            ok = ok && (kb_shr(w, 16u + sp) & 1u);
        }
        const uint32_t q4 = __byte_perm(a, 0u, 0x4442u);  // byte 2: 4 * process (one PRMT)
        const uint32_t nsa = wns + q4;
        const uint32_t nsw = kb_lds32(nsa);
        const int nq = (int)(nsw & 0xffffu);
        int lo = (int)(nsw >> 16);
        const uint32_t plane = wb + __byte_perm(a, 0u, 0x4441u) * pl2;
        const uint32_t ea = plane + ca2;
        const uint32_t e = kb_lds16(ea);
        const uint32_t wina = win + 2u * q4;
        const uint32_t tag = (a >> 11) & 0xe000u;  // member << 13
        const bool is_add = (int)a < 0;
        const uint32_t t = e ^ tag;  // < 0x2000: registered by this member, and t is its position
        const bool del_go = valid && !is_add && t < 0x2000u;
        const bool add_try = ok && is_add;
        const bool add_go = add_try && e == 0u && nq < C;
        if (add_try && !add_go) bad |= 1u;
        const uint32_t lb = q4 * capH;  // byte offset of the process' list
        // the element a del moves into the freed position: the list's last one
        uint32_t last = kb_lds16(wina + 2u * (uint32_t)((nq - 1) & 3));
        if (del_go && lo >= nq) {
            const int c0 = (nq - 1) & ~3;
            last = kb_gen_refill(p1 + lb + 2u * (uint32_t)c0, wina, nq);
            lo = c0;
        }
        const bool move = del_go && (int)t < nq;
        // operands computed unconditionally, five predicated stores: no branch in the round's tail
        const bool wr = add_go || move, upd = add_go || del_go;
        const int idx = add_go ? nq : (int)t - 1;
        const uint32_t val = add_go ? (ca2 >> 1) : last;
        const bool wr_win = wr && (add_go || idx >= lo);
        const int nq2 = add_go ? nq + 1 : nq - 1;
        const int lo2 = add_go ? max(lo, nq - 3) : lo;
        const uint32_t ent = add_go ? (tag | (uint32_t)(nq + 1)) : 0u;
        if (wr) *list_at(lb + 2u * (uint32_t)idx) = (uint16_t)val;
        if (wr_win) kb_sts16(wina + 2u * (uint32_t)(idx & 3), val);
        if (move) kb_sts16(plane + 2u * last, e);
        if (upd) kb_sts16(ea, ent);
        if (upd) kb_sts32(nsa, (uint32_t)nq2 | ((uint32_t)lo2 << 16));
        __syncwarp();
    }
};

// One Philox4x32-10 block -> this lane's uniforms: slot 0: a = -log(ran_time) with ran_time in (0,1],
// b = ran_proc; slot 1: a = ran_site in [0,1).
__device__ __forceinline__ void kb_gen_uniforms(const unsigned long long st, const uint32_t replica_id,
                                                const uint32_t slot, const uint32_t k0, const uint32_t k1,
                                                double& a, double& b) {
    uint32_t rnd[4];
    kb_philox4x32_10((uint32_t)st, (uint32_t)(st >> 32), replica_id, slot, k0, k1, rnd);
    const double u0 = (double)(((((uint64_t)rnd[1] << 32) | rnd[0]) >> 11) + (uint64_t)(slot ^ 1u)) * 0x1.0p-53;
    const double u1 = (double)((((uint64_t)rnd[3] << 32) | rnd[2]) >> 11) * 0x1.0p-53;
    a = slot ? u0 : -log(u0);
    b = u1;
}

// serial float64 chain over the packed non-zero products: the lane adds T entries ending at `top`, the
// first (T - own count) of them leading zeros (adding 0.0 is exact, so this is base.mpy:615-618's
// left-to-right recurrence restricted to the non-zero terms)
template <int T>
__device__ __forceinline__ double kb_gen_chain(const double* top) {
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < T; ++t) acc = __dadd_rn(acc, top[t - T]);
    return acc;
}

template <class M>
__global__ void __launch_bounds__(M::MAX_THREADS) kb_gen_kernel(const KbGenParams prm) {
    extern __shared__ __align__(128) unsigned char kb_sm[];
    constexpr int P = M::P, LPR = M::LPR, G = 32 / LPR;
    constexpr int PPL = (P + LPR - 1) / LPR;  // processes per lane: lane sl owns PPL*sl .. PPL*sl + PPL-1
    constexpr int NPAD = PPL * LPR;           // nr_of_sites words per replica in shared memory (zero beyond P)
    constexpr int BATCH = LPR / 2;            // steps of uniforms generated at once
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPR, sl = lane % LPR;
    unsigned gmask = KB_FULL;
    if (LPR < 32) gmask = ((1u << (LPR & 31)) - 1u) << (g * LPR);
    const unsigned ltg = gmask & kb_lanemask_lt();
    {
        const uint4* src = reinterpret_cast<const uint4*>(prm.tab);
        uint4* dst = reinterpret_cast<uint4*>(kb_sm);
        for (int i = threadIdx.x; i < prm.tab_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    // no block-wide barrier below this line: every warp is an independent worker
    unsigned char* const wblk = kb_sm + prm.tab_bytes + (size_t)warp * prm.warp_bytes;
    unsigned char* const rb = wblk + (size_t)g * prm.rep_bytes;  // this group's replica block
    uint8_t* const latp = rb + prm.sm_lat;
    uint32_t* const nS = reinterpret_cast<uint32_t*>(rb + prm.sm_ns);
    double* const Zp = reinterpret_cast<double*>(rb + prm.sm_prod) + KB_GEN_ZEROS;
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(kb_sm + prm.mbar_off) + warp;
    if (prm.use_bulk) {
        if (lane == 0) {
            kb_mbar_init(mbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    uint32_t mbar_phase = 0;
    KbGenCtx<M> c;
    c.wb = kb_smem_addr(rb);
    c.lat = c.wb + (uint32_t)prm.sm_lat;
    c.wns = c.wb + (uint32_t)prm.sm_ns;
    c.win = c.wb + (uint32_t)prm.sm_win;
    c.tab0 = kb_smem_addr(kb_sm);
    c.tabA = c.tab0 + 4u * (uint32_t)sl;
    c.tabB = c.tab0 + (uint32_t)M::OFF_B + (uint32_t)(4 * M::BW) * (uint32_t)sl;
    c.nbT = c.tab0 + (uint32_t)prm.nbt_off;
    c.sl = sl; c.C = prm.ncells; c.capH = (uint32_t)prm.cap >> 1; c.pl2 = 2u * (uint32_t)prm.ncells;
    const int cap = prm.cap;

    for (;;) {  // ---- persistent worker loop: one (epoch, team of G replicas) item per iteration -----------
        int item = 0;
        if (lane == 0) item = atomicAdd(prm.work_counter, 1);
        item = __shfl_sync(KB_FULL, item, 0);
        if (item >= prm.n_items) break;
        const int epoch = item / prm.n_teams;
        const int team = item - epoch * prm.n_teams;
        long long my_steps = prm.chunk;
        if ((long long)epoch * prm.chunk + my_steps > prm.nsteps) my_steps = prm.nsteps - (long long)epoch * prm.chunk;
        if (epoch > 0) {
            if (lane == 0) {
                while (kb_ld_acquire(prm.done + team) < epoch) __nanosleep(128);
            }
            __syncwarp();
            __threadfence();
            kb_fence_proxy_async_all();
        }
        const int rep = team * G + g;
        const bool have = rep < prm.R;
        const int repc = have ? rep : prm.R - 1;  // addresses of an absent group stay valid, nothing is stored
        unsigned char* const g_img = prm.image + (size_t)repc * prm.img_bytes;
        uint8_t* const g_lat = prm.lattice + (size_t)repc * prm.lat_stride;
        int32_t* const g_ns = prm.nsites + (size_t)repc * P;
        {
            unsigned long long p1v = reinterpret_cast<unsigned long long>(g_img);
            asm volatile("" : "+l"(p1v));  // keep the list base in registers
            c.p1 = reinterpret_cast<unsigned char*>(p1v);
        }
        // ---- stage plane 2 and the lattice into shared memory (TMA bulk copy, mbarrier completion) --------
        if (prm.use_bulk) {
            if (lane == 0) {
                const int n_here = min(G, prm.R - team * G);
                kb_mbar_expect_tx(mbar, (uint32_t)(n_here * (prm.stage_bytes + prm.lat_stride)));
                for (int gg = 0; gg < n_here; ++gg) {
                    const size_t r2 = (size_t)(team * G + gg);
                    unsigned char* dst = wblk + (size_t)gg * prm.rep_bytes;
                    kb_bulk_g2s(dst, prm.image + r2 * prm.img_bytes + prm.stage_off, (uint32_t)prm.stage_bytes, mbar);
                    kb_bulk_g2s(dst + prm.sm_lat, prm.lattice + r2 * prm.lat_stride, (uint32_t)prm.lat_stride, mbar);
                }
            }
            kb_mbar_wait(mbar, mbar_phase);
            mbar_phase ^= 1u;
        } else if (have) {
            const uint4* s1 = reinterpret_cast<const uint4*>(g_img + prm.stage_off);
            uint4* d1 = reinterpret_cast<uint4*>(rb);
            for (int i = sl; i < prm.stage_bytes / 16; i += LPR) d1[i] = s1[i];
            const uint4* sl4 = reinterpret_cast<const uint4*>(g_lat);
            uint4* dl = reinterpret_cast<uint4*>(latp);
            for (int i = sl; i < prm.lat_stride / 16; i += LPR) dl[i] = sl4[i];
        }
        for (int i = sl; i < NPAD; i += LPR) {
            const uint32_t n = (have && i < P) ? (uint32_t)g_ns[i] : 0u;
            nS[i] = n | (n << 16);  // window empty: lo = nr_of_sites
        }
        for (int i = sl; i < KB_GEN_ZEROS; i += LPR) Zp[i - KB_GEN_ZEROS] = 0.0;
        __syncwarp();

        // ---- per-lane process registers -----------------------------------------------------------------
        double rate[PPL], integ[PPL];
        uint32_t cnt[PPL];
        bool has[PPL];
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int q = PPL * sl + j;
            has[j] = q < P;
            const bool ld = has[j] && have;
            rate[j] = ld ? prm.rates[(size_t)repc * P + q] : 0.0;
            integ[j] = ld ? prm.integ[(size_t)repc * P + q] : 0.0;
            cnt[j] = 0;
        }
        KbScalars* const scp = prm.sc + repc;
        double kmc_time = scp->kmc_time, kmc_dt = scp->kmc_time_step;
        int status = scp->status;
        const long long step0 = scp->kmc_step;
        const uint32_t replica_id = scp->replica;
        const uint32_t k0 = (uint32_t)scp->seed, k1 = (uint32_t)(scp->seed >> 32);
        bool act = have && status == KB_OK;  // this group is stepping
        int nst = 0;                          // steps this replica has done in this item
        c.bad = 0;
        double rng_a = 0.0, rng_b = 0.0;  // even lanes: (-log(ran_time), ran_proc); odd lanes: (ran_site, -)

        const int n_it = (int)my_steps;
        bool alive = __any_sync(KB_FULL, act);  // warp-uniform: some group of this warp is still stepping
        for (int it = 0; it < n_it && alive; ++it) {
            bool stopped = false;               // this group stopped in this step (dead-lock): re-vote at the end
            const int sub = it & (BATCH - 1);
            if (sub == 0) {
                // BATCH steps of uniforms at once: lane sl serves step kmc_step + sl/2, Philox slot sl&1
                kb_gen_uniforms((unsigned long long)(step0 + it) + (unsigned)(sl >> 1), replica_id, (uint32_t)(sl & 1),
                                k0, k1, rng_a, rng_b);
            }
            const double neg_log_u = __shfl_sync(KB_FULL, rng_a, 2 * sub, LPR);
            const double ran_proc = __shfl_sync(KB_FULL, rng_b, 2 * sub, LPR);
            const double ran_site = __shfl_sync(KB_FULL, rng_a, 2 * sub + 1, LPR);

            // -- update_accum_rate over the packed non-zero products
            double pr[PPL];
            if (PPL == 2) {
                const uint2 nn = reinterpret_cast<const uint2*>(nS)[sl];
                pr[0] = __dmul_rn((double)(int)(nn.x & 0xffffu), rate[0]);
                pr[PPL - 1] = __dmul_rn((double)(int)(nn.y & 0xffffu), rate[PPL - 1]);
            } else if (PPL == 4) {
                const uint4 nn = reinterpret_cast<const uint4*>(nS)[sl];
                const uint32_t v[4] = {nn.x, nn.y, nn.z, nn.w};
#pragma unroll
                for (int j = 0; j < PPL; ++j) pr[j] = __dmul_rn((double)(int)(v[j & 3] & 0xffffu), rate[j]);
            } else {
#pragma unroll
                for (int j = 0; j < PPL; ++j) pr[j] = __dmul_rn((double)(int)(nS[PPL * sl + j] & 0xffffu), rate[j]);
            }
            int below = 0, ctot = 0;
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const unsigned nz = __ballot_sync(KB_FULL, pr[j] != 0.0);
                below += __popc(nz & ltg);
                ctot += __popc(nz & gmask);
            }
            {
                int pos = below;
#pragma unroll
                for (int j = 0; j < PPL; ++j)
                    if (pr[j] != 0.0) { Zp[pos] = pr[j]; ++pos; }
            }
            const int cmax = G == 1 ? ctot : (int)__reduce_max_sync(KB_FULL, (unsigned)ctot);
            __syncwarp();
            const int cnt0 = below + (pr[0] != 0.0 ? 1 : 0);  // non-zero products up to and including process PPL*sl
            double acc[PPL];
            {
                const double* top = Zp + cnt0;
                if (P <= 4 || cmax <= 4) acc[0] = kb_gen_chain<4>(top);
                else if (P <= 8 || cmax <= 8) acc[0] = kb_gen_chain<8>(top);
                else if (P <= 16 || cmax <= 16) acc[0] = kb_gen_chain<16>(top);
                else {
                    double a0 = 0.0;
                    for (int t = 0; t < cmax; ++t) {
                        const double x = Zp[t];
                        a0 = __dadd_rn(a0, t < cnt0 ? x : 0.0);  // + 0.0 behind the lane's own terms: exact
                    }
                    acc[0] = a0;
                }
            }
#pragma unroll
            for (int j = 1; j < PPL; ++j) acc[j] = __dadd_rn(acc[j - 1], pr[j]);
            const double total = __shfl_sync(KB_FULL, acc[PPL - 1], LPR - 1, LPR);  // the last lane has added every product
            if (act && !(total > 0.0)) { status = KB_DEADLOCK; act = false; stopped = true; }

            const bool act_clk = act;  // the clock advances for this step (update_clocks below, behind the site read)

            // -- determine_procsite: first process whose accumulated rate exceeds ran_proc*total
            double value = __dmul_rn(ran_proc, total);
            // value >= accum(P) (ran_proc*total rounded up to the total): the reference's search ends on the last
            // entry and then walks left over entries that are >= their right neighbour (base.mpy:1316-1326),
            // i.e. it returns the first process whose accumulated rate equals the total.  Searching for the
            // largest double below the total finds exactly that process: no second pass, no branch.
            if (!(value < total)) value = __longlong_as_double(__double_as_longlong(total) - 1LL);
            int pidx = 0;
#pragma unroll
            for (int j = 0; j < PPL; ++j) pidx += __popc(__ballot_sync(KB_FULL, has[j] && !(value < acc[j])) & gmask);
            if (!act || pidx >= P) pidx = 0;
            const int nsel = (int)(nS[pidx] & 0xffffu);
            if (act && nsel <= 0) { status = KB_DEADLOCK; act = false; stopped = true; }  // the clock has advanced: the step counts
            int k = (int)__dadd_rn(1.0, __dmul_rn(ran_site, (double)nsel));
            k = max(min(k, nsel), 1);

            // -- run_proc_nr(pidx + 1, site): the event's row of the descriptor table
            const uint4 ev = kb_ldc128(c.tab0 + (uint32_t)M::OFF_EV + 16u * (uint32_t)pidx);
            // determine_procsite's site read: avail_sites(proc, k, 1) (base.mpy:1110-1113)
            uint32_t cell = *c.list_at(2u * (uint32_t)(pidx * cap + k - 1));
            // -- update_clocks / update_integ_rate, issued behind the site read: they do not depend on it and
            // the division's latency disappears in the L2 round trip (+4 %)
            asm volatile("" : "+r"(cell));
            {   // a stopped group divides by its zero total and discards the result: no branch
                const double dt = neg_log_u / total;
                kmc_dt = act_clk ? dt : kmc_dt;
                kmc_time = act_clk ? __dadd_rn(kmc_time, dt) : kmc_time;
#pragma unroll
                for (int j = 0; j < PPL; ++j) integ[j] = act_clk ? __dadd_rn(integ[j], __dmul_rn(pr[j], dt)) : integ[j];
                nst += act_clk ? 1 : 0;
            }
            // this lane's replace_species call (lanes 0..3 of the group), 0 = none
            const uint32_t wr = kb_ldc32(c.tab0 + (uint32_t)M::OFF_WR + 16u * (uint32_t)pidx + 4u * (uint32_t)(sl & 3));
            const int nr = act ? (int)(ev.y & 0xffu) : 0;
            const int nr_max = G == 1 ? nr : (int)__reduce_max_sync(KB_FULL, (unsigned)nr);
            uint32_t rda = c.tab0 + ev.x;
            uint32_t d = nr > 0 ? kb_ldc32(rda) : 0u;
            uint32_t a = c.ldA(d & 0xffffu);
            KbGenOpB b = c.ldB(d & 0xffffu);
            // increment_procstat (base.mpy:1010-1023): the lane that owns the process counts the event
#pragma unroll
            for (int j = 0; j < PPL; ++j) cnt[j] += (act && pidx == PPL * sl + j) ? 1u : 0u;
            c.nbrow = c.nbT + cell * (2 * M::NOFF);
            // replace_species(site, old, new) (base.mpy:1187-1231): wr = column | (site-1) << 8 | old << 16 | new << 24
            {   // every lane computes (an idle lane reads column 0 / site 1: valid), two predicated tails
                const bool wv = act && sl < 4 && wr != 0u;
                const uint32_t c2 = kb_ldc16(c.nbrow + __byte_perm(wr, 0u, 0x4440u));
                const uint32_t p = c.lat + c.lat_index(c2) + (__byte_perm(wr, 0u, 0x4441u) & 0x7fu);
                const bool match = kb_lds8(p) == __byte_perm(wr, 0u, 0x4442u);
                if (wv && match) kb_sts8(p, wr >> 24);
                if (wv && !match) c.bad |= 2u << sl;
            }
            // rounds: the next round's operands are requested before the current round runs
            for (int r = 0; r < nr_max; ++r) {
                rda += 4u;
                const uint32_t dn = (r + 1 < nr) ? kb_ldc32(rda) : 0u;
                const uint32_t an = c.ldA(dn & 0xffffu);
                const KbGenOpB bn = c.ldB(dn & 0xffffu);
                c.round((int)((d >> 16) & 0xffu), a, b);
                d = dn; a = an; b = bn;
            }
            __syncwarp();  // lattice writes of an event without ops must be visible to the next step

            if (__any_sync(KB_FULL, c.bad != 0 || stopped)) {
                uint32_t gb = 0;  // OR of the group's flags
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    if (__ballot_sync(KB_FULL, (c.bad >> i) & 1u) & gmask) gb |= 1u << i;
                if (gb >> 1) {
                    // error tuple of the first failing replace_species call (old, new, found, site, step); the
                    // site was left untouched (base.mpy:1205-1228, KMC_Model.post_mortem)
                    const int wi = __ffs(gb >> 1) - 1;
                    const uint32_t w = prm.writes[4 * pidx + wi];
                    const uint32_t off = w & 255u, n = (w >> 8) & 255u;
                    const uint32_t c2 = kb_ldc16(c.nbrow + 2 * off);
                    const int idx = (int)c.lat_index(c2) + (int)n - 1;
                    if (sl == 0) {
                        scp->err[0] = (int)((w >> 16) & 255u); scp->err[1] = (int)((w >> 24) & 255u);
                        scp->err[2] = latp[idx]; scp->err[3] = idx + 1; scp->err[4] = (int)(step0 + nst);
                    }
                    status = KB_SPECIES_MISMATCH;
                    act = false;
                } else if (gb) {
                    status = KB_CAPACITY;
                    act = false;
                }
                c.bad = 0;
                alive = __any_sync(KB_FULL, act);
            }
        }
        const long long kmc_step = step0 + nst;

        // ---- write back ---------------------------------------------------------------------------------
        __syncwarp();
        if (prm.use_bulk) {
            kb_fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                const int n_here = min(G, prm.R - team * G);
                for (int gg = 0; gg < n_here; ++gg) {
                    const size_t r2 = (size_t)(team * G + gg);
                    unsigned char* src = wblk + (size_t)gg * prm.rep_bytes;
                    kb_bulk_s2g(prm.image + r2 * prm.img_bytes + prm.stage_off, src, (uint32_t)prm.stage_bytes);
                    kb_bulk_s2g(prm.lattice + r2 * prm.lat_stride, src + prm.sm_lat, (uint32_t)prm.lat_stride);
                }
                kb_bulk_commit_wait();
                kb_fence_proxy_async_all();
            }
        } else if (have) {
            uint4* s1 = reinterpret_cast<uint4*>(g_img + prm.stage_off);
            const uint4* d1 = reinterpret_cast<const uint4*>(rb);
            for (int i = sl; i < prm.stage_bytes / 16; i += LPR) s1[i] = d1[i];
            uint4* sl4 = reinterpret_cast<uint4*>(g_lat);
            const uint4* dl = reinterpret_cast<const uint4*>(latp);
            for (int i = sl; i < prm.lat_stride / 16; i += LPR) sl4[i] = dl[i];
        }
        if (have) {
            for (int i = sl; i < P; i += LPR) g_ns[i] = (int32_t)(nS[i] & 0xffffu);
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                if (has[j]) {
                    const size_t o = (size_t)rep * P + (size_t)(PPL * sl + j);
                    prm.integ[o] = integ[j];
                    prm.procstat[o] += cnt[j];
                }
            }
            if (sl == 0) {
                scp->kmc_time = kmc_time; scp->kmc_time_step = kmc_dt; scp->kmc_step = kmc_step; scp->status = status;
            }
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) kb_st_release(prm.done + team, epoch + 1);
        __syncwarp();  // the shared-memory blocks are reused by the next item
    }
}
#endif  // __CUDACC__

// ---- host side: layout, tables, launch (compiled into the generated module) ---------------------------
static inline uint64_t kb_gen_fnv1a(const void* data, size_t bytes) {
    const unsigned char* p = (const unsigned char*)data;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static inline int kb_gen_align(int x, int a) { return (x + a - 1) / a * a; }

// image + shared-memory layout for one geometry; warps per CTA and CTAs per SM chosen from the device's
// limits so that as many replicas as possible are resident.  returns 0, or a negative reason code
static inline int kb_gen_make_plan(const KbGenInfo& gi, const int size[3], int R, int device, int regs,
                                   int max_threads, KbGenPlan* pl) {
    memset(pl, 0, sizeof *pl);
    long long cells = 1;
    for (int a = 0; a < 3; ++a) { pl->size[a] = a < gi.dim ? size[a] : 1; cells *= pl->size[a]; }
    if (cells > (long long)KB_POS_MASK) return -1;  // position field of a class entry
    for (int i = 0; i < gi.n_off; ++i)
        for (int a = 0; a < gi.dim; ++a) {
            int o = gi.offsets[3 * i + a];
            if (o < 0) o = -o;
            if (2 * o >= pl->size[a]) return -2;  // two offsets would alias under the periodic wrap
        }
    pl->ncells = (int)cells; pl->R = R; pl->device = device;
    pl->cap = kb_gen_align(pl->ncells, 8);
    pl->off_p2 = kb_gen_align(gi.n_proc * pl->cap * 2, 16);
    pl->img_bytes = kb_gen_align(pl->off_p2 + gi.n_classes * pl->ncells * 2, 16);
    pl->stage_off = pl->off_p2;
    pl->stage_bytes = pl->img_bytes - pl->stage_off;
    pl->lat_stride = kb_gen_align(pl->ncells * gi.spuck, 16);
    if ((long long)gi.n_proc * pl->cap * 2 > 0x7fffffffLL) return -1;
    const int lpr = gi.lpr, groups = 32 / lpr;
    const int ppl = (gi.n_proc + lpr - 1) / lpr, npad = ppl * lpr;
    pl->lpr = lpr;
    pl->mbar_off = gi.ops_bytes;            // one mbarrier per warp (32 x 8 bytes)
    pl->nbt_off = gi.ops_bytes + 256;
    pl->tab_bytes = kb_gen_align(pl->nbt_off + pl->ncells * gi.n_off * 2, 16);
    pl->sm_lat = pl->stage_bytes;
    pl->sm_ns = kb_gen_align(pl->sm_lat + pl->lat_stride, 16);
    pl->sm_win = pl->sm_ns + 4 * npad;
    pl->sm_prod = pl->sm_win + 8 * gi.n_proc;
    pl->rep_bytes = kb_gen_align(pl->sm_prod + 8 * (KB_GEN_ZEROS + gi.n_proc), 16);
    pl->warp_bytes = groups * pl->rep_bytes;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return -3; }
    pl->sm_count = prop.multiProcessorCount;
    pl->regs = regs;
    const int max_smem = (int)prop.sharedMemPerBlockOptin;
    const int per_sm = (int)prop.sharedMemPerMultiprocessor;
    const int regs8 = (regs + 7) & ~7;
    const int reg_warps = prop.regsPerMultiprocessor / (32 * regs8);
    int best = 0;
    for (int w = 1; w <= 32 && w * 32 <= max_threads; ++w) {
        const int w4 = (w + 3) & ~3;
        if (w4 > reg_warps) break;
        const int bytes = pl->tab_bytes + w * pl->warp_bytes;
        if (bytes > max_smem) break;
        int n = per_sm / (bytes + 1024);
        if (n * w > 64) n = 64 / w;
        if (n * w4 > reg_warps) n = reg_warps / w4;
        if (n < 1) continue;
        if (n * w > best || (n * w == best && n < pl->ctas_per_sm)) {
            best = n * w; pl->wpc = w; pl->ctas_per_sm = n; pl->smem_bytes = bytes;
        }
    }
    if (!best) return -4;
    pl->replicas_per_cta = pl->wpc * groups;
    return 0;
}

// operand table for one geometry (layout: KbGenInfo), then the neighbour table nbT[cell][offset] = 2 * cell
// index of (cell + offset) under the periodic wrap
static inline void kb_gen_fill_tables(const KbGenInfo& gi, const KbGenPlan& pl, uint32_t* out) {
    memset(out, 0, (size_t)pl.tab_bytes);
    unsigned char* base = (unsigned char*)out;
    // idle lanes of a round read operands past its last op: every op slot (the padding too) holds a valid word
    const int n_slots = gi.n_ops + 32;
    for (int i = 0; i < n_slots; ++i) {
        uint32_t* b = (uint32_t*)(base + gi.off_b) + gi.bw * i;
        for (int j = 0; j < gi.bw; ++j) b[j] = 0xffff0000u;
    }
    uint32_t* rdw = (uint32_t*)(base + gi.off_rd);
    for (int r = 0; r < gi.n_rounds; ++r) {
        const KbGenRoundDesc& rd = gi.rounds[r];
        rdw[r] = (uint32_t)rd.first_op | ((uint32_t)rd.count << 16) | ((uint32_t)rd.kind << 24);
        for (int i = 0; i < rd.count; ++i) {
            const KbGenOpDesc& op = gi.ops[rd.first_op + i];
            ((uint32_t*)base)[rd.first_op + i] = (2u * op.aoff) | ((uint32_t)op.cls << 8) | ((4u * op.q) << 16) |
                                                 ((uint32_t)op.member << 24) | (op.is_add ? KB_GEN_KIND_ADD : 0u);
            uint32_t* b = (uint32_t*)(base + gi.off_b) + gi.bw * (rd.first_op + i);
            for (int j = 0; j < gi.bw; ++j) {
                if (j < op.ncond) b[j] = 2u * op.coff[j] | (((uint32_t)op.cn[j] - 1u) << 8) | ((uint32_t)op.cmask[j] << 16);
                else b[j] = 0xffff0000u;  // always true: any species of site 1 of the event's own cell
            }
        }
    }
    rdw[gi.n_rounds] = 0;
    uint32_t* evw = (uint32_t*)(base + gi.off_ev);
    uint32_t* wrw = (uint32_t*)(base + gi.off_wr);
    for (int p = 0; p < gi.n_proc; ++p) {
        int nw = 0;
        for (int i = 0; i < 4; ++i) {
            const uint32_t x = gi.writes[4 * p + i];  // off id | n << 8 | old << 16 | new << 24, 0 = none
            wrw[4 * p + i] = 0;
            if (!x) continue;
            // bit 15 marks a real write: a check-only call on species 0 of site 1 of column 0 would read as 0
            wrw[4 * p + nw++] = (2u * (x & 255u)) | ((((x >> 8) & 255u) - 1u) << 8) | (x & 0xffff0000u) | 0x8000u;
        }
        evw[4 * p] = (uint32_t)gi.off_rd + 4u * (uint32_t)gi.events[p].first_round;
        evw[4 * p + 1] = (uint32_t)gi.events[p].n_rounds | ((uint32_t)nw << 8);
        evw[4 * p + 2] = 0; evw[4 * p + 3] = 0;
    }
    uint16_t* nbt = (uint16_t*)(base + pl.nbt_off);
    const int Lx = pl.size[0], Ly = pl.size[1], Lz = pl.size[2];
    for (int cell = 0; cell < pl.ncells; ++cell) {
        const int z = cell / (Lx * Ly), y = (cell / Lx) % Ly, x = cell % Lx;
        for (int o = 0; o < gi.n_off; ++o) {
            const int xx = ((x + gi.offsets[3 * o]) % Lx + Lx) % Lx;
            const int yy = ((y + gi.offsets[3 * o + 1]) % Ly + Ly) % Ly;
            const int zz = ((z + gi.offsets[3 * o + 2]) % Lz + Lz) % Lz;
            nbt[cell * gi.n_off + o] = (uint16_t)(2 * (xx + Lx * (yy + Ly * zz)));
        }
    }
}

// launch geometry: persistent CTAs, the launch cut into epochs so that the tail is a fraction of an epoch
// (persistent scheduling).  sched: [1 + teams] ints, zeroed here on the stream.
template <class K>
static inline int kb_gen_do_launch(K kernel, const KbGenInfo& gi, const KbGenPlan& pl, KbGenParams p, int* sched,
                                   int wpc_override, int epochs_override, cudaStream_t stream) {
    const int groups = 32 / gi.lpr;
    int wpc = pl.wpc;
    if (wpc_override > 0 && wpc_override <= pl.wpc) wpc = wpc_override;
    const int smem = pl.tab_bytes + wpc * pl.warp_bytes;
    const int teams = (p.R + groups - 1) / groups;
    int blocks = (teams + wpc - 1) / wpc;
    const int resident = pl.sm_count * pl.ctas_per_sm;
    if (blocks > resident) blocks = resident;
    const long long slots = (long long)blocks * wpc;
    const long long n = p.nsteps;
    long long epochs = 1;
    if (teams > slots) {
        epochs = (24 * slots + teams - 1) / teams;
        const long long max_epochs = n / 256 > 0 ? n / 256 : 1;
        if (epochs > max_epochs) epochs = max_epochs;
        if (epochs > 64) epochs = 64;
        if (epochs < 1) epochs = 1;
    }
    if (epochs_override > 0) epochs = epochs_override;
    if ((n + epochs - 1) / epochs > 0x40000000LL) epochs = (n + 0x3fffffffLL) / 0x40000000LL;
    p.chunk = (n + epochs - 1) / epochs;
    epochs = (n + p.chunk - 1) / p.chunk;
    if (epochs * (long long)teams > 0x7fffffffLL) return -1;
    p.n_teams = teams;
    p.n_items = (int)(epochs * teams);
    p.work_counter = sched;
    p.done = sched + 1;
    cudaError_t e = cudaMemsetAsync(sched, 0, ((size_t)teams + 1) * sizeof(int), stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    kernel<<<blocks, wpc * 32, smem, stream>>>(p);
    return (int)cudaGetLastError();
}

// The C entry points of a generated module (dlopen'ed by libkmos_b200.so, kmos_b200_batch_attach_proclist).
#define KB_GEN_MODULE(M, INFO)                                                                                  \
    extern "C" const KbGenInfo* kmos_b200_gen_info(void) { return &(INFO); }                                   \
    extern "C" int kmos_b200_gen_plan(const int32_t size[3], int32_t R, int32_t device, KbGenPlan* out) {      \
        cudaFuncAttributes fa;                                                                                  \
        if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return -3; }                           \
        if (cudaFuncGetAttributes(&fa, kb_gen_kernel<M>) != cudaSuccess) { cudaGetLastError(); return -3; }    \
        return kb_gen_make_plan(INFO, size, R, device, fa.numRegs, fa.maxThreadsPerBlock, out);                \
    }                                                                                                           \
    extern "C" int kmos_b200_gen_build_tables(const KbGenPlan* pl, uint32_t* out) {                            \
        kb_gen_fill_tables(INFO, *pl, out);                                                                     \
        return 0;                                                                                               \
    }                                                                                                           \
    extern "C" int kmos_b200_gen_launch(const KbGenParams* p, const KbGenPlan* pl, int* sched, int wpc,        \
                                        int epochs, void* stream) {                                            \
        return kb_gen_do_launch(kb_gen_kernel<M>, INFO, *pl, *p, sched, wpc, epochs, (cudaStream_t)stream);    \
    }
