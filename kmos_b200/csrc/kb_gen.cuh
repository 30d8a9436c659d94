// kb_gen.cuh -- skeleton of the exporter-generated per-model step kernel (sm_100a).
//
// kmos is a code generator: run_proc_nr and the put_/take_ routines it calls exist only as generated,
// model-specific straight-line code (kmos/io/__init__.py:305-465 write_proclist_run_proc_nr_smart,
// :2219-2409 write_proclist_put_take, :2568-2655 _write_optimal_iftree).  kmos_b200/codegen.py emits the
// CUDA counterpart, proclist_<model>.cu: one `case` per process with the event's lattice writes as
// immediates and its guarded del_proc / if-tree add_proc calls as unrolled rounds -- which lanes delete,
// which add, how many probes a round has and where its operands sit are compile-time facts of the case.
// This header is the model-independent part that file instantiates:
//
//   do_kmc_steps loop        proclist_generic_subroutines.mpy:1-44     kb_gen_kernel<M>
//   update_accum_rate        base.mpy:603-623    packed non-zero products, serial float64 chain per lane
//   update_clocks            base.mpy:1123-1161
//   update_integ_rate        base.mpy:626-645
//   determine_procsite       base.mpy:1075-1120, interval_search_real :1234-1338 as a warp ballot
//   add_proc / del_proc      base.mpy:211-302    KbGenCtx::round<>
//   replace_species          base.mpy:1187-1231  KbGenCtx::write<>
//
// One warp steps one replica.  Plane 2 of avail_sites (one uint16 entry per exclusivity class and cell, see
// kmos_b200/devtables.py) and the lattice stay in shared memory for the whole work item; plane 1 (one list per
// process, always growing upwards) stays in HBM/L2.  The CTA's warps share the model's operand table: per
// op one uint4 with ready-made byte offsets (neighbour-table column, class plane, 4 * process) and the member
// tag, plus one packed word per if-tree probe, specialised on the host for the lattice geometry
// (kb_gen_build_tables).
//
// Two code styles, chosen by the generator from the size of the model's event code:
//   unrolled  every process is its own straight-line case (what kmos does in Fortran).  Fastest while the
//             whole kernel stays inside the SM's instruction cache (mini_101, AB: 2-3 k SASS instructions).
//   compact   RuO2's 36 cases unroll to 12 k instructions = 196 KB of SASS; with 22 warps of a CTA in
//             different cases the instruction cache thrashes (ncu r2: 6.5 cycles of no_instruction stall per
//             issued instruction, IPC 1.3).  Here the generator emits only the round bodies the model uses
//             (dispatch() = a switch over its (del, add, probes) variants) and the events become rows of a
//             descriptor table: writes and rounds are read, not decoded field by field.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "kb_smem.cuh"

#define KB_GEN_ABI 3
#define KB_GEN_MAX_COND 4
#define KB_GEN_KIND_ADD 0x80000000u
#define KB_GEN_TABLE_PAD 1024  // idle lanes of a round read up to 32 entries past its last op

// ---- static description of a generated module (host side) ----------------------------------------------
struct KbGenOpDesc {  // one list operation, in table order
    uint8_t is_add, q, cls, member, aoff, ncond;
    uint8_t coff[KB_GEN_MAX_COND], cn[KB_GEN_MAX_COND];
    uint16_t cmask[KB_GEN_MAX_COND];
};
struct KbGenRoundDesc {  // one round: `count` ops from `first_op`, at most `nc` probes each
    int32_t first_op, count, nc, kind;  // kind: index of the round body in the model's dispatch()
};
struct KbGenEventDesc {  // one process: its rounds
    int32_t first_round, n_rounds;
};
struct KbGenInfo {
    int32_t abi, n_proc, n_species, spuck, dim, n_off, n_classes, n_ops, n_rounds;
    // operand table: [A: uint4 per op][B: bw probe words per op][round words][event rows][neighbour table]
    int32_t bw, off_b, off_rd, off_ev, ops_bytes, max_threads, compact;
    uint64_t model_hash;  // FNV-1a of the int32 model blob the code was generated from
    const char* name;
    const KbGenOpDesc* ops;
    const KbGenRoundDesc* rounds;
    const KbGenEventDesc* events;
    const int8_t* offsets;     // [n_off][3]
    const uint32_t* writes;    // [n_proc][4]: off_id | n<<8 | old<<16 | new<<24, 0 = none (error reporting)
    const uint8_t* proc_cls;   // [n_proc] exclusivity class
    const uint8_t* proc_member;  // [n_proc] member tag (1..7)
};

// geometry-dependent layout, computed by kmos_b200_gen_plan
struct KbGenPlan {
    int32_t size[3], ncells, R, device;
    // compact image in HBM: [plane 1: n_proc lists of `cap` uint16][plane 2: n_classes x ncells uint16]
    int32_t cap, off_p2, img_bytes, stage_off, stage_bytes, lat_stride;
    // shared memory: [table][per-warp blocks]; offsets inside a block
    int32_t tab_bytes, nbt_off, rep_bytes, sm_lat, sm_ns, sm_prod, sm_rng, sm_mbar;
    int32_t wpc, ctas_per_sm, smem_bytes, regs, sm_count;
};

struct KbGenParams {
    const uint32_t* tab;  // device copy of the table built by kmos_b200_gen_build_tables
    int tab_bytes, nbt_off;
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
    int* work_counter;
    int* done;
    int n_items;
    long long chunk;
    int rep_bytes, sm_lat, sm_ns, sm_prod, sm_rng, sm_mbar;
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

__device__ __forceinline__ uint2 kb_ldc64(uint32_t a) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

struct KbGenOpB {  // up to 4 if-tree probes: column (byte 0), site (byte 1), species mask (high half)
    uint32_t c[KB_GEN_MAX_COND];
};

// the replica's mutable shared-memory state (class planes, nr_of_sites, lattice) by 32-bit shared address
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
__device__ __forceinline__ void kb_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void kb_sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void kb_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Everything an event's generated code works with.  M: the generated model traits.
template <class M>
struct KbGenCtx {
    uint32_t wb;         // shared address of this warp's block: class planes at 0
    uint32_t wns;        // ... of its nr_of_sites array
    uint32_t lat;        // ... of its lattice copy
    unsigned char* p1;   // the replica's lists in HBM/L2 (list of process q at byte q*cap*2)
    uint32_t tab0;       // shared address of the operand table
    uint32_t tabA;       // ... + lane*16: this lane's uint4 of a round whose first op sits at offset 0
    uint32_t tabB;       // ... + OFF_B + lane*4*BW: this lane's probe words
    uint32_t nbT;        // shared address of the neighbour table: row = cell, column = offset, value = 2*cell'
    uint32_t nbrow;      // row of the selected cell
    int lane, C, cap;
    uint32_t capH;       // cap / 2: list byte offset of process q = (4*q) * capH
    uint32_t cell2;      // 2 * selected cell
    uint32_t bad;        // bit 0: capacity, bit 1+i: write i found another species
    uint32_t cntA, cntB; // events of this lane's processes in the current work item

    __device__ __forceinline__ uint16_t* list_at(uint32_t byte_off) const {
        return reinterpret_cast<uint16_t*>(p1 + byte_off);
    }
    // determine_procsite's site read: avail_sites(proc, k, 1) (base.mpy:1110-1113)
    template <int Q>
    __device__ __forceinline__ void select(int k) {
        int cp = cap;
        asm volatile("" : "+r"(cp));  // keeps the compiler from hoisting every case's list base in front of the switch
        const uint32_t cell = *list_at(2u * (uint32_t)(Q * cp + k - 1));
        // increment_procstat (base.mpy:1010-1023): the lane that owns process Q counts the event
        if (M::P > 32) {
            if (Q & 1) cntB += (lane == Q / 2);
            else cntA += (lane == Q / 2);
        } else {
            cntA += (lane == Q);
        }
        cell2 = 2u * cell;
        nbrow = nbT + cell * (2 * M::NOFF);
    }
    __device__ __forceinline__ void select_rt(const int q, const int k) {  // the same, process number in a register
        const uint32_t cell = *list_at(2u * (uint32_t)(q * cap + k - 1));
        if (M::P > 32) {
            const uint32_t hit = lane == (q >> 1);
            cntB += hit & (uint32_t)q;
            cntA += hit & ~(uint32_t)q;
        } else {
            cntA += (lane == q);
        }
        cell2 = 2u * cell;
        nbrow = nbT + cell * (2 * M::NOFF);
    }
    __device__ __forceinline__ uint32_t lat_index(uint32_t c2) const {
        return (M::SPUCK % 2 == 0) ? c2 * (M::SPUCK / 2) : (c2 * M::SPUCK) >> 1;
    }
    // replace_species(site, old, new) (base.mpy:1187-1231); every lane executes the same write
    template <int I, int OFF, int N, int OLD, int NEW>
    __device__ __forceinline__ void write() {
        const uint32_t c2 = OFF == 0 ? cell2 : kb_ldc16(nbrow + 2 * OFF);
        const uint32_t p = lat + lat_index(c2) + (N - 1);
        if (kb_lds8(p) == OLD) kb_sts8(p, NEW);
        else bad |= 2u << I;
    }
    // the same from an event row: w = column | (site - 1) << 8 | old << 16 | new << 24
    __device__ __forceinline__ void write_rt(const int i, const uint32_t w) {
        const uint32_t c2 = kb_ldc16(nbrow + (w & 0xffu));
        const uint32_t p = lat + lat_index(c2) + ((w >> 8) & 0xffu);
        if (kb_lds8(p) == ((w >> 16) & 0xffu)) kb_sts8(p, w >> 24);
        else bad |= 2u << i;
    }
    // operands of the round whose first op has index I (lane l: op I + l)
    __device__ __forceinline__ uint4 ldA(const uint32_t i) const { return kb_ldc128(tabA + 16u * i); }
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

    // One round: lane l < count executes op l of the round.  a = (neighbour column, class plane, 4 * process,
    // member << 13 | (add ? 1 << 31 : 0)), the first two as byte offsets.
    //   guarded del_proc (base.mpy:211-265): registered iff the class entry carries this op's member tag
    //   add_proc (base.mpy:268-302) after the if-tree probes of its leaf (io/__init__.py:2568-2655)
    template <bool HAS_DEL, bool HAS_ADD, int NC>
    __device__ __forceinline__ void round(const int count, const uint4 a, const KbGenOpB& b) {
        const bool valid = lane < count;
        const uint32_t ca2 = kb_ldc16(nbrow + a.x);
        bool ok = valid;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const uint32_t w = b.c[j];
            const uint32_t cc2 = kb_ldc16(nbrow + (w & 0xffu));
            const uint32_t sp = kb_lds8(lat + ((w >> 8) & 0xffu) + lat_index(cc2));
            ok = ok && (kb_shr(w, 16u + sp) & 1u);
        }
        const uint32_t nsa = wns + a.z;
        const int nq = (int)kb_lds32(nsa);
        const uint32_t plane = wb + a.y;
        const uint32_t ea = plane + ca2;
        const uint32_t e = kb_lds16(ea);
        const uint32_t lb = a.z * capH;  // byte offset of the process' list
        const uint32_t kk = a.w;
        const bool is_add = HAS_ADD && (!HAS_DEL || (int)kk < 0);
        if (HAS_DEL && !HAS_ADD) {
            const uint32_t t = (e ^ kk) & 0xffffu;  // < 0x2000: registered, and t is its position
            const bool go = valid && t < 0x2000u;
            // the element a del would move, requested before anything depends on it
            const uint32_t last = *list_at(lb + 2u * (uint32_t)((valid && nq > 0) ? nq - 1 : 0));
            if (go) {
                if ((int)t < nq) {
                    *list_at(lb + 2u * (t - 1u)) = (uint16_t)last;
                    kb_sts16(plane + 2u * last, e);
                }
                kb_sts16(ea, 0u);
                kb_sts32(nsa, (uint32_t)(nq - 1));
            }
        } else if (HAS_ADD && !HAS_DEL) {
            const bool go = ok && e == 0 && nq < C;
            if (ok && !go) bad |= 1u;
            if (go) {
                *list_at(lb + 2u * (uint32_t)nq) = (uint16_t)(ca2 >> 1);
                kb_sts16(ea, (kk & 0xffffu) | (uint32_t)(nq + 1));
                kb_sts32(nsa, (uint32_t)(nq + 1));
            }
        } else {
            const uint32_t t = (e ^ kk) & 0xffffu;
            const bool want_last = valid && !is_add && nq > 0;
            const uint32_t last = *list_at(lb + 2u * (uint32_t)(want_last ? nq - 1 : 0));
            const bool add_try = ok && is_add;
            const bool add_go = add_try && e == 0 && nq < C;
            const bool del_go = valid && !is_add && t < 0x2000u;
            const bool move = del_go && (int)t < nq;
            if (add_try && !add_go) bad |= 1u;
            if (add_go || move) *list_at(lb + 2u * (uint32_t)(add_go ? nq : (int)t - 1)) = (uint16_t)(add_go ? (ca2 >> 1) : last);
            if (move) kb_sts16(plane + 2u * last, e);
            if (add_go || del_go) {
                kb_sts16(ea, add_go ? ((kk & 0xffffu) | (uint32_t)(nq + 1)) : 0u);
                kb_sts32(nsa, (uint32_t)(add_go ? nq + 1 : nq - 1));
            }
        }
        __syncwarp();
    }
};

// compact style: the event as a row of the descriptor table.  Row (uint4): byte offset of its first round
// word, n_rounds | n_writes << 8, writes 0 and 1 (writes 2 and 3 in a second array).  Round word: first op |
// count << 16 | kind << 24; the next round's operands are requested before the current round runs (the word
// behind an event's last round is the next event's first or the table's terminator, so the read is harmless).
template <class M>
__device__ __forceinline__ void kb_gen_run_compact(KbGenCtx<M>& c, const int pidx, const int k) {
    const uint4 ev = kb_ldc128(c.tab0 + (uint32_t)M::OFF_EV + 16u * (uint32_t)pidx);
    c.select_rt(pidx, k);
    uint32_t rda = c.tab0 + ev.x;
    uint32_t d = kb_ldc32(rda);
    uint4 a = c.ldA(d & 0xffffu);
    KbGenOpB b = c.ldB(d & 0xffffu);
    const int nw = (int)(ev.y >> 8), nr = (int)(ev.y & 0xffu);
    if (nw > 0) c.write_rt(0, ev.z);
    if (nw > 1) c.write_rt(1, ev.w);
    if (nw > 2) {
        const uint2 e2 = kb_ldc64(c.tab0 + (uint32_t)M::OFF_EV + 16u * (uint32_t)M::P + 8u * (uint32_t)pidx);
        c.write_rt(2, e2.x);
        if (nw > 3) c.write_rt(3, e2.y);
    }
    if (nr == 0) __syncwarp();
    for (int r = 0; r < nr; ++r) {
        rda += 4u;
        const uint32_t dn = kb_ldc32(rda);
        const uint4 an = c.ldA(dn & 0xffffu);
        const KbGenOpB bn = c.ldB(dn & 0xffffu);
        M::dispatch(c, d >> 24, (int)((d >> 16) & 0xffu), a, b);
        d = dn; a = an; b = bn;
    }
}

// serial float64 chain over the packed non-zero products: lane adds `tier` entries ending at `top`, the
// first (tier - own count) of them leading zeros (adding 0.0 is exact, so this is base.mpy:615-618's
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
    constexpr int P = M::P;
    constexpr int PPL = P > 32 ? 2 : 1;
    constexpr int NPAD = 32 * PPL;  // nr_of_sites entries per replica in shared memory (zero beyond P)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const uint4* src = reinterpret_cast<const uint4*>(prm.tab);
        uint4* dst = reinterpret_cast<uint4*>(kb_sm);
        for (int i = threadIdx.x; i < prm.tab_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    // no block-wide barrier below this line: every warp is an independent worker
    unsigned char* const wb = kb_sm + prm.tab_bytes + (size_t)warp * prm.rep_bytes;
    int32_t* const nS = reinterpret_cast<int32_t*>(wb + prm.sm_ns);
    double* const Zp = reinterpret_cast<double*>(wb + prm.sm_prod) + 32;  // 32 leading zeros, then NPAD packed products
    double* const rngS = reinterpret_cast<double*>(wb + prm.sm_rng);      // 16 steps x (-log ran_time, ran_proc, ran_site)
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(wb + prm.sm_mbar);
    if (prm.use_bulk) {
        if (lane == 0) {
            kb_mbar_init(mbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    uint32_t mbar_phase = 0;
    KbGenCtx<M> c;
    c.wb = kb_smem_addr(wb);
    c.lat = kb_smem_addr(wb) + (uint32_t)prm.sm_lat;
    uint8_t* const latp = wb + prm.sm_lat;
    c.wns = kb_smem_addr(wb) + (uint32_t)prm.sm_ns;
    c.tab0 = kb_smem_addr(kb_sm);
    c.tabA = kb_smem_addr(kb_sm) + 16u * (uint32_t)lane;
    c.tabB = kb_smem_addr(kb_sm) + (uint32_t)M::OFF_B + (uint32_t)(4 * M::BW) * (uint32_t)lane;
    c.nbT = kb_smem_addr(kb_sm) + (uint32_t)prm.nbt_off;
    c.lane = lane; c.C = prm.ncells; c.cap = prm.cap; c.capH = (uint32_t)prm.cap >> 1;
    const unsigned lt = kb_lanemask_lt();

    for (;;) {  // ---- persistent worker loop: one (epoch, replica) item per iteration ---------------------
        int item = 0;
        if (lane == 0) item = atomicAdd(prm.work_counter, 1);
        item = __shfl_sync(KB_FULL, item, 0);
        if (item >= prm.n_items) break;
        const int epoch = item / prm.R;
        const int rep = item - epoch * prm.R;
        long long my_steps = prm.chunk;
        if ((long long)epoch * prm.chunk + my_steps > prm.nsteps) my_steps = prm.nsteps - (long long)epoch * prm.chunk;
        if (epoch > 0) {
            if (lane == 0) {
                while (kb_ld_acquire(prm.done + rep) < epoch) __nanosleep(128);
            }
            __syncwarp();
            __threadfence();
            kb_fence_proxy_async_all();
        }
        unsigned char* const g_img = prm.image + (size_t)rep * prm.img_bytes;
        unsigned char* const g_stage = g_img + prm.stage_off;
        uint8_t* const g_lat = prm.lattice + (size_t)rep * prm.lat_stride;
        int32_t* const g_ns = prm.nsites + (size_t)rep * P;
        {
            unsigned long long p1v = reinterpret_cast<unsigned long long>(g_img);
            asm volatile("" : "+l"(p1v));  // keep the list base in registers (see kb_smem.cuh)
            c.p1 = reinterpret_cast<unsigned char*>(p1v);
        }
        // ---- stage plane 2 and the lattice into shared memory (TMA bulk copy, mbarrier completion) --------
        if (prm.use_bulk) {
            if (lane == 0) {
                kb_mbar_expect_tx(mbar, (uint32_t)(prm.stage_bytes + prm.lat_stride));
                kb_bulk_g2s(wb, g_stage, (uint32_t)prm.stage_bytes, mbar);
                kb_bulk_g2s(latp, g_lat, (uint32_t)prm.lat_stride, mbar);
            }
            kb_mbar_wait(mbar, mbar_phase);
            mbar_phase ^= 1u;
        } else {
            const uint4* s1 = reinterpret_cast<const uint4*>(g_stage);
            uint4* d1 = reinterpret_cast<uint4*>(wb);
            for (int i = lane; i < prm.stage_bytes / 16; i += 32) d1[i] = s1[i];
            const uint4* sl = reinterpret_cast<const uint4*>(g_lat);
            uint4* dl = reinterpret_cast<uint4*>(latp);
            for (int i = lane; i < prm.lat_stride / 16; i += 32) dl[i] = sl[i];
        }
        for (int i = lane; i < NPAD; i += 32) nS[i] = i < P ? g_ns[i] : 0;
        for (int i = lane; i < 32; i += 32) Zp[i - 32] = 0.0;
        __syncwarp();

        // ---- per-lane process registers: lane L owns processes PPL*L .. PPL*L + PPL-1 ---------------------
        const int qa = PPL * lane, qb = PPL * lane + 1;
        const bool hasA = qa < P, hasB = PPL == 2 && qb < P;
        const double rateA = hasA ? prm.rates[(size_t)rep * P + qa] : 0.0;
        const double rateB = hasB ? prm.rates[(size_t)rep * P + qb] : 0.0;
        double integA = hasA ? prm.integ[(size_t)rep * P + qa] : 0.0;
        double integB = hasB ? prm.integ[(size_t)rep * P + qb] : 0.0;
        KbScalars* const scp = prm.sc + rep;
        double kmc_time = scp->kmc_time, kmc_dt = scp->kmc_time_step;
        int status = scp->status;
        const long long step0 = scp->kmc_step;
        const uint32_t replica_id = scp->replica;
        const uint32_t k0 = (uint32_t)scp->seed, k1 = (uint32_t)(scp->seed >> 32);
        c.cntA = 0; c.cntB = 0;
        c.bad = 0;
        int pidx = 0;

        const int n_it = (int)my_steps;
        int it = 0;
        for (; it < n_it && status == KB_OK; ++it) {
            const int sub = it & (KB_RNG_BATCH - 1);
            if (sub == 0) {
                // 16 steps of uniforms at once: lane l serves step kmc_step + l/2, Philox slot l&1
                const unsigned long long st = (unsigned long long)(step0 + it) + (unsigned)(lane >> 1);
                uint32_t rnd[4];
                kb_philox4x32_10((uint32_t)st, (uint32_t)(st >> 32), replica_id, (uint32_t)(lane & 1), k0, k1, rnd);
                const double u0 = (double)(((((uint64_t)rnd[1] << 32) | rnd[0]) >> 11) + (uint64_t)((lane & 1) ^ 1)) * 0x1.0p-53;
                const double u1 = (double)((((uint64_t)rnd[3] << 32) | rnd[2]) >> 11) * 0x1.0p-53;
                double* dst = rngS + 3 * (lane >> 1);
                if (lane & 1) {
                    dst[2] = u0;  // ran_site in [0,1)
                } else {
                    dst[0] = -log(u0);  // ran_time in (0,1]
                    dst[1] = u1;        // ran_proc
                }
                __syncwarp();
            }
            const double neg_log_u = rngS[3 * sub], ran_proc = rngS[3 * sub + 1], ran_site = rngS[3 * sub + 2];

            // -- update_accum_rate over the packed non-zero products
            int nA, nB = 0;
            if (PPL == 2) {
                const int2 nn = reinterpret_cast<const int2*>(nS)[lane];
                nA = nn.x; nB = nn.y;
            } else {
                nA = nS[lane];
            }
            const double prA = __dmul_rn((double)nA, rateA);
            const double prB = PPL == 2 ? __dmul_rn((double)nB, rateB) : 0.0;
            const unsigned nzA = __ballot_sync(KB_FULL, prA != 0.0);
            const unsigned nzB = PPL == 2 ? __ballot_sync(KB_FULL, prB != 0.0) : 0u;
            const int posA = __popc(nzA & lt) + (PPL == 2 ? __popc(nzB & lt) : 0);
            const int cA = posA + (prA != 0.0 ? 1 : 0);  // non-zero products up to and including process A
            if (prA != 0.0) Zp[posA] = prA;
            if (PPL == 2 && prB != 0.0) Zp[cA] = prB;
            __syncwarp();
            const int c_tot = __popc(nzA) + (PPL == 2 ? __popc(nzB) : 0);
            const double* top = Zp + cA;
            double accA;
            if (P <= 4 || c_tot <= 4) accA = kb_gen_chain<4>(top);
            else if (P <= 8 || c_tot <= 8) accA = kb_gen_chain<8>(top);
            else if (P <= 16 || c_tot <= 16) accA = kb_gen_chain<16>(top);
            else if (NPAD == 32 || c_tot <= 32) accA = kb_gen_chain<32>(top);
            else {
                accA = 0.0;
                for (int t = 0; t < cA; ++t) accA = __dadd_rn(accA, Zp[t]);
            }
            const double accB = PPL == 2 ? __dadd_rn(accA, prB) : accA;
            const double total = __shfl_sync(KB_FULL, accB, 31);  // lane 31 has added every product
            if (!(total > 0.0)) { status = KB_DEADLOCK; break; }

            // -- update_clocks / update_integ_rate
            kmc_dt = neg_log_u / total;
            kmc_time = __dadd_rn(kmc_time, kmc_dt);
            integA = __dadd_rn(integA, __dmul_rn(prA, kmc_dt));
            if (PPL == 2) integB = __dadd_rn(integB, __dmul_rn(prB, kmc_dt));

            // -- determine_procsite: first process whose accumulated rate exceeds ran_proc*total
            const double value = __dmul_rn(ran_proc, total);
            const unsigned leA = __ballot_sync(KB_FULL, hasA && !(value < accA));
            const unsigned leB = PPL == 2 ? __ballot_sync(KB_FULL, hasB && !(value < accB)) : 0u;
            pidx = __popc(leA) + __popc(leB);
            if (pidx >= P) {
                // value >= accum(P): the reference's search ends on the last entry and then walks left over
                // entries that are >= their right neighbour (base.mpy:1316-1326)
                const unsigned geA = __ballot_sync(KB_FULL, hasA && accA >= total);
                const unsigned geB = PPL == 2 ? __ballot_sync(KB_FULL, hasB && accB >= total) : 0u;
                pidx = P - (__popc(geA) + __popc(geB));
            }
            const int nsel = nS[pidx];
            if (nsel <= 0) { status = KB_DEADLOCK; ++it; break; }  // the clock has advanced: the step counts
            int k = (int)__dadd_rn(1.0, __dmul_rn(ran_site, (double)nsel));
            k = min(k, nsel);

            // -- run_proc_nr(pidx + 1, site): the generated per-process code
            M::run_event(c, pidx, k);

            if (__any_sync(KB_FULL, c.bad != 0)) {
                const unsigned all = __reduce_or_sync(KB_FULL, c.bad);
                if (all >> 1) {
                    // error tuple of the first failing replace_species call (old, new, found, site, step); the
                    // site was left untouched (base.mpy:1205-1228, KMC_Model.post_mortem)
                    const int wi = __ffs(all >> 1) - 1;
                    const uint32_t w = prm.writes[4 * pidx + wi];
                    const uint32_t off = w & 255u, n = (w >> 8) & 255u;
                    const uint32_t c2 = kb_ldc16(c.nbrow + 2 * off);
                    const int idx = (int)c.lat_index(c2) + (int)n - 1;
                    if (lane == 0) {
                        scp->err[0] = (int)((w >> 16) & 255u); scp->err[1] = (int)((w >> 24) & 255u);
                        scp->err[2] = latp[idx]; scp->err[3] = idx + 1; scp->err[4] = (int)(step0 + it + 1);
                    }
                    status = KB_SPECIES_MISMATCH;
                } else {
                    status = KB_CAPACITY;
                }
            }
        }
        const long long kmc_step = step0 + it;

        // ---- write back ---------------------------------------------------------------------------------
        __syncwarp();
        if (prm.use_bulk) {
            kb_fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                kb_bulk_s2g(g_stage, wb, (uint32_t)prm.stage_bytes);
                kb_bulk_s2g(g_lat, latp, (uint32_t)prm.lat_stride);
                kb_bulk_commit_wait();
                kb_fence_proxy_async_all();
            }
        } else {
            uint4* s1 = reinterpret_cast<uint4*>(g_stage);
            const uint4* d1 = reinterpret_cast<const uint4*>(wb);
            for (int i = lane; i < prm.stage_bytes / 16; i += 32) s1[i] = d1[i];
            uint4* sl = reinterpret_cast<uint4*>(g_lat);
            const uint4* dl = reinterpret_cast<const uint4*>(latp);
            for (int i = lane; i < prm.lat_stride / 16; i += 32) sl[i] = dl[i];
        }
        for (int i = lane; i < P; i += 32) g_ns[i] = nS[i];
        if (hasA) { prm.integ[(size_t)rep * P + qa] = integA; prm.procstat[(size_t)rep * P + qa] += c.cntA; }
        if (hasB) { prm.integ[(size_t)rep * P + qb] = integB; prm.procstat[(size_t)rep * P + qb] += c.cntB; }
        if (lane == 0) {
            scp->kmc_time = kmc_time; scp->kmc_time_step = kmc_dt; scp->kmc_step = kmc_step; scp->status = status;
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) kb_st_release(prm.done + rep, epoch + 1);
        __syncwarp();  // the shared-memory block is reused by the next item
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

// image + shared-memory layout for one geometry; wpc/ctas chosen from the device's limits.
// returns 0, or a negative reason code
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
    const int npad = gi.n_proc > 32 ? 64 : 32;
    pl->nbt_off = gi.ops_bytes;
    pl->tab_bytes = kb_gen_align(gi.ops_bytes + pl->ncells * gi.n_off * 2, 128);
    pl->sm_lat = pl->stage_bytes;
    pl->sm_ns = kb_gen_align(pl->sm_lat + pl->lat_stride, 16);
    pl->sm_prod = kb_gen_align(pl->sm_ns + 4 * npad, 16);
    pl->sm_rng = pl->sm_prod + 8 * (32 + npad);
    pl->sm_mbar = pl->sm_rng + 8 * 3 * KB_RNG_BATCH;
    pl->rep_bytes = kb_gen_align(pl->sm_mbar + 16, 128);
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
        const int bytes = pl->tab_bytes + w * pl->rep_bytes;
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
    return 0;
}

// operand table for one geometry (layout: KbGenInfo), then the neighbour table nbT[cell][offset] = 2 * cell
// index of (cell + offset) under the periodic wrap
static inline void kb_gen_fill_tables(const KbGenInfo& gi, const KbGenPlan& pl, uint32_t* out) {
    memset(out, 0, (size_t)pl.tab_bytes);
    unsigned char* base = (unsigned char*)out;
    uint32_t* rdw = (uint32_t*)(base + gi.off_rd);
    for (int r = 0; r < gi.n_rounds; ++r) {
        const KbGenRoundDesc& rd = gi.rounds[r];
        rdw[r] = (uint32_t)rd.first_op | ((uint32_t)rd.count << 16) | ((uint32_t)rd.kind << 24);
        for (int i = 0; i < rd.count; ++i) {
            const KbGenOpDesc& op = gi.ops[rd.first_op + i];
            uint32_t* a = (uint32_t*)base + 4 * (rd.first_op + i);
            a[0] = 2u * op.aoff;
            a[1] = (uint32_t)op.cls * (uint32_t)pl.ncells * 2u;
            a[2] = 4u * op.q;
            a[3] = ((uint32_t)op.member << KB_POS_BITS) | (op.is_add ? KB_GEN_KIND_ADD : 0u);
            uint32_t* b = (uint32_t*)(base + gi.off_b) + gi.bw * (rd.first_op + i);
            for (int j = 0; j < gi.bw; ++j) {
                if (j < op.ncond) b[j] = 2u * op.coff[j] | (((uint32_t)op.cn[j] - 1u) << 8) | ((uint32_t)op.cmask[j] << 16);
                else b[j] = 0xffff0000u;  // always true: any species of site 1 of the event's own cell
            }
        }
    }
    rdw[gi.n_rounds] = 0;  // terminator: the prefetch behind the last event's last round reads op 0
    uint32_t* evw = (uint32_t*)(base + gi.off_ev);
    uint32_t* ev2 = evw + 4 * gi.n_proc;
    for (int p = 0; p < gi.n_proc; ++p) {
        uint32_t w[4];
        int nw = 0;
        for (int i = 0; i < 4; ++i) {
            const uint32_t x = gi.writes[4 * p + i];  // off id | n << 8 | old << 16 | new << 24, 0 = none
            w[i] = 0;
            if (!x) continue;
            w[nw++] = (2u * (x & 255u)) | ((((x >> 8) & 255u) - 1u) << 8) | (x & 0xffff0000u);
        }
        evw[4 * p] = (uint32_t)gi.off_rd + 4u * (uint32_t)gi.events[p].first_round;
        evw[4 * p + 1] = (uint32_t)gi.events[p].n_rounds | ((uint32_t)nw << 8);
        evw[4 * p + 2] = w[0]; evw[4 * p + 3] = w[1];
        ev2[2 * p] = w[2]; ev2[2 * p + 1] = w[3];
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
    extern "C" int kmos_b200_gen_launch(const KbGenParams* p, int blocks, int threads, int smem, void* stream) { \
        cudaError_t e = cudaFuncSetAttribute(kb_gen_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return (int)e;                                                                    \
        kb_gen_kernel<M><<<blocks, threads, smem, (cudaStream_t)stream>>>(*p);                                  \
        return (int)cudaGetLastError();                                                                         \
    }
