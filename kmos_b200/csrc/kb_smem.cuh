// kb_smem.cuh -- shared-memory kMC step kernel for sm_100a: one warp steps one replica; the replica's
// lattice and (compacted) avail-site tables stay in shared memory for the whole launch, TMA bulk-staged
// from/to HBM; the warps of a CTA share one copy of the model's per-event lane tables.
//
// Reference loop restated per step (kmos/fortran_src/proclist_generic_subroutines.mpy:22-42):
//   random_number x3            per-replica Philox4x32-10 counter stream, generated 16 steps at a time
//                               (lane l: step base+l/2, slot l&1) together with -log(ran_time)
//   update_accum_rate           base.mpy:603-623   products nr_of_sites(q)*rates(q) go through shared
//                                                  memory; every lane runs the serial left-to-right
//                                                  float64 recurrence up to its own process (bit parity)
//   update_clocks               base.mpy:1123-1161
//   update_integ_rate           base.mpy:626-645   lane-parallel over processes
//   determine_procsite          base.mpy:1075-1120 + interval_search_real :1234-1338 as a warp ballot
//   run_proc_nr                 generated put_/take_ routines (kmos/io/__init__.py:305-465, 2219-2409)
//                               flattened by kmos_b200/devtables.py into rounds of list ops, one per lane;
//                               add_proc/del_proc (base.mpy:211-302) on the compact storage below
//
// Compact avail-site storage (see devtables.py): plane 1 per arena (two mutually exclusive processes share
// `cap` uint16 slots, one list from the left, one from the right), plane 2 per exclusivity class and cell
// (uint16 = member << 13 | position).
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

#define KB_POS_BITS 13
#define KB_POS_MASK 0x1FFFu
#define KB_RNG_BATCH 16

struct KbSmemParams {
    // model / geometry
    const int32_t* dev;  // lane tables in global memory: SEC_DEVICE with ops specialised for this geometry
    int dev_words;
    int nbt_bytes;       // bytes of the per-CTA neighbour table (0: compute neighbours arithmetically)
    int n_proc, spuck, dim;
    int size[3];
    int ncells, volume;
    int n_arenas, n_classes, cap;  // cap = ncells + spare slots per arena
    uint32_t magic_x, magic_xy;    // umulhi(c, magic) == c / Lx  (resp. c / (Lx*Ly)) for all c < ncells
    // batch arrays (global)
    uint8_t* lattice;   // [R][lat_stride]
    int32_t* nsites;    // [R][P]
    uint16_t* image;    // [R][img_bytes/2]  compact planes: arenas then class entries
    uint16_t* p1;       // canonical planes [R][plane_bytes/2] (pack / unpack kernels only)
    uint16_t* p2;
    const double* rates;
    double* integ;
    int64_t* procstat;
    KbScalars* sc;
    int R;
    long long nsteps;
    // persistent scheduling: work item i = (epoch i / R, replica i % R); an epoch is `chunk` steps.  Warps
    // fetch items from `work_counter`; `done[r]` counts the finished epochs of replica r in this launch.
    int* work_counter;
    int* done;
    int n_items;
    long long chunk;
    // shared-memory layout (bytes)
    int tab_bytes, rep_bytes;
    int off_hi, off_p2;                                   // offsets inside the compact image
    int sm_p2, sm_lat, sm_ns, sm_prod, sm_mbar;           // offsets inside a replica's shared-memory block
    int stage_off, stage_bytes;                           // part of the image that is staged into smem
    int split;        // 1: plane 1 stored as low bytes + a bitmap of bit 8 (ncells <= 512), 0: uint16
    int p1_global;    // 1: plane 1 (the lists) stays in HBM/L2, only plane 2 + lattice live in shared memory
    int lat_stride;   // bytes, multiple of 16
    int img_bytes;    // bytes of the compact image (both planes), multiple of 16
    int plane_bytes;  // bytes of one canonical plane, multiple of 16
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
__device__ __forceinline__ void kb_fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int kb_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void kb_st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int kb_s8(uint32_t w, int shift) { return (int)(int8_t)((w >> shift) & 255u); }

__device__ __forceinline__ unsigned kb_lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ unsigned kb_lanemask_le() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// slot index of position k (0-based) of a list inside its arena
__device__ __forceinline__ int kb_slot(int arena, int dir, int cap, int k) {
    return arena * cap + (dir ? cap - 1 - k : k);
}

// Plane-1 accessors.  SPLIT: a cell index < 512 is stored as its low byte plus one bit in a bitmap
// (1.125 B per slot instead of 2): for 20x20 lattices this is what lets 13 instead of 9 replicas share an SM.
// Slots of one byte of the bitmap belong to one list (the two lists of an arena stay >= 8 slots apart, see
// the spare-slot rule in plan_smem), and ops on one list are serialised by the round schedule.
template <bool SPLIT>
__device__ __forceinline__ int kb_p1_get(const unsigned char* p1, const unsigned char* hi, int slot) {
    if (SPLIT) {
        const uint32_t lo = p1[slot];
        const uint32_t hb = hi[slot >> 3];
        return (int)(lo | (((hb >> (slot & 7)) & 1u) << 8));
    }
    return (int)reinterpret_cast<const uint16_t*>(p1)[slot];
}
template <bool SPLIT>
__device__ __forceinline__ void kb_p1_set(unsigned char* p1, unsigned char* hi, int slot, int cell) {
    if (SPLIT) {
        p1[slot] = (unsigned char)cell;
        const uint32_t bit = 1u << (slot & 7);
        const uint32_t hb = hi[slot >> 3];
        hi[slot >> 3] = (unsigned char)((cell & 256) ? (hb | bit) : (hb & ~bit));
    } else {
        reinterpret_cast<uint16_t*>(p1)[slot] = (uint16_t)cell;
    }
}

// ---------------------------------------------------------------------------------------------------
// canonical <-> compact conversion (one warp per replica, runs between engine switches only)
// ---------------------------------------------------------------------------------------------------
__global__ void kb_pack_kernel(const KbSmemParams prm) {
    const int lane = threadIdx.x & 31;
    const int rep = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (rep >= prm.R) return;
    const int P = prm.n_proc, C = prm.ncells, cap = prm.cap;
    const int32_t* procinfo = prm.dev + prm.dev[9];
    unsigned char* img = reinterpret_cast<unsigned char*>(prm.image) + (size_t)rep * prm.img_bytes;
    uint16_t* cp2 = reinterpret_cast<uint16_t*>(img + prm.off_p2);
    const uint16_t* g1 = prm.p1 + (size_t)rep * (prm.plane_bytes / 2);
    const int32_t* ns = prm.nsites + (size_t)rep * P;
    for (int i = lane; i < prm.img_bytes / 4; i += 32) reinterpret_cast<uint32_t*>(img)[i] = 0;
    __syncwarp();
    for (int q = 0; q < P; ++q) {
        const uint32_t pi = (uint32_t)procinfo[q];
        const int arena = pi & 63, dir = (pi >> 6) & 1, cls = (pi >> 7) & 31, member = (pi >> 12) & 7;
        const int n = ns[q];
        for (int k = lane; k < n; k += 32) {
            const int cell = g1[(size_t)q * C + k];
            if (!prm.split) reinterpret_cast<uint16_t*>(img)[kb_slot(arena, dir, cap, k)] = (uint16_t)cell;
            else img[kb_slot(arena, dir, cap, k)] = (unsigned char)cell;
            cp2[(size_t)cls * C + cell] = (uint16_t)((member << KB_POS_BITS) | (k + 1));
        }
        if (prm.split) {  // bit 8 of every slot, one bitmap byte per lane and pass (no shared bytes)
            __syncwarp();
            unsigned char* hi = img + prm.off_hi;
            const int first = dir ? cap - n : 0;  // slots [first, first+n) of this arena
            for (int b = (first >> 3) + lane; b <= ((first + n - 1) >> 3) && n > 0; b += 32) {
                uint32_t byte = hi[(arena * cap >> 3) + b];
                for (int j = 0; j < 8; ++j) {
                    const int sl = b * 8 + j;
                    if (sl < first || sl >= first + n) continue;
                    const int k = dir ? cap - 1 - sl : sl;
                    const int cell = g1[(size_t)q * C + k];
                    byte = (cell & 256) ? (byte | (1u << j)) : (byte & ~(1u << j));
                }
                hi[(arena * cap >> 3) + b] = (unsigned char)byte;
            }
            __syncwarp();
        }
    }
}

__global__ void kb_unpack_kernel(const KbSmemParams prm) {
    const int lane = threadIdx.x & 31;
    const int rep = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (rep >= prm.R) return;
    const int P = prm.n_proc, C = prm.ncells, cap = prm.cap;
    const int32_t* procinfo = prm.dev + prm.dev[9];
    const unsigned char* img = reinterpret_cast<const unsigned char*>(prm.image) + (size_t)rep * prm.img_bytes;
    uint16_t* g1 = prm.p1 + (size_t)rep * (prm.plane_bytes / 2);
    uint16_t* g2 = prm.p2 + (size_t)rep * (prm.plane_bytes / 2);
    const int32_t* ns = prm.nsites + (size_t)rep * P;
    for (int i = lane; i < prm.plane_bytes / 2; i += 32) { g1[i] = 0; g2[i] = 0; }
    __syncwarp();
    for (int q = 0; q < P; ++q) {
        const uint32_t pi = (uint32_t)procinfo[q];
        const int arena = pi & 63, dir = (pi >> 6) & 1;
        const int n = ns[q];
        for (int k = lane; k < n; k += 32) {
            const int sl = kb_slot(arena, dir, cap, k);
            const int cell = prm.split ? kb_p1_get<true>(img, img + prm.off_hi, sl) : kb_p1_get<false>(img, img, sl);
            g1[(size_t)q * C + k] = (uint16_t)cell;
            g2[(size_t)q * C + cell] = (uint16_t)(k + 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// step kernel
// ---------------------------------------------------------------------------------------------------
struct KbCellCtx {
    int Lx, Ly, Lz, dim;
    uint32_t magic_x, magic_xy;
    int x, y, z;
    __device__ __forceinline__ void decode(int cell) {
        if (dim >= 3) {
            z = (int)__umulhi((uint32_t)cell, magic_xy);
            cell -= z * Lx * Ly;
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
    // cell index of (x+dx, y+dy, z+dz) with periodic wrap; |d*| < L* is checked at batch creation
    __device__ __forceinline__ int cell_at(uint32_t packed) const {
        int xx = x + kb_s8(packed, 0);
        xx += (xx < 0) ? Lx : 0;
        xx -= (xx >= Lx) ? Lx : 0;
        int c = xx;
        if (dim >= 2) {
            int yy = y + kb_s8(packed, 8);
            yy += (yy < 0) ? Ly : 0;
            yy -= (yy >= Ly) ? Ly : 0;
            c += Lx * yy;
        }
        if (dim >= 3) {
            int zz = z + kb_s8(packed, 16);
            zz += (zz < 0) ? Lz : 0;
            zz -= (zz >= Lz) ? Lz : 0;
            c += Lx * Ly * zz;
        }
        return c;
    }
};

// PPL: processes per lane (1: P <= 32, 2: P <= 64);  NCOND: largest number of dynamic probes of an add
// Specialised op (host: specialise_tables in kmos_b200.cu): 2 + NCOND words
//   w0 = kind | ncond<<1 | dir<<4 | (4*q)<<8 | off_id<<16 | member<<24   (byte-aligned: one PRMT per field)
//   w1 = class base (cls * ncells) | first slot of the list (arena*cap, or arena*cap + cap-1 if it grows down) << 16
//   then ncond probe words off_id | n<<5 | mask<<8
template <int PPL, int NCOND, bool SPLIT, bool P1G, bool NBT>
__global__ void __launch_bounds__(768) kb_smem_kernel(const KbSmemParams prm) {
    extern __shared__ __align__(128) unsigned char kb_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    int32_t* tab = reinterpret_cast<int32_t*>(kb_smem);
    for (int i = threadIdx.x; i < prm.dev_words; i += blockDim.x) tab[i] = prm.dev[i];
    __syncthreads();
    const uint4* events = reinterpret_cast<const uint4*>(tab + tab[3]);
    const uint32_t* ops = reinterpret_cast<const uint32_t*>(tab + tab[4]);
    const uint32_t* offsets = reinterpret_cast<const uint32_t*>(tab + tab[7]);
    const uint32_t* procinfo = reinterpret_cast<const uint32_t*>(tab + tab[9]);
    const int n_off = tab[8];
    constexpr int STRIDE = 2 + NCOND;

    KbCellCtx cc;
    cc.Lx = prm.size[0]; cc.Ly = prm.size[1]; cc.Lz = prm.size[2]; cc.dim = prm.dim;
    cc.magic_x = prm.magic_x; cc.magic_xy = prm.magic_xy;
    // neighbour table shared by the CTA's replicas: nbT[cell][o] = cell index of (cell + offset o)
    uint16_t* nbT = reinterpret_cast<uint16_t*>(kb_smem + prm.tab_bytes);
    if (NBT) {
        for (int idx = threadIdx.x; idx < prm.ncells * n_off; idx += blockDim.x) {
            const int c = idx / n_off;
            cc.decode(c);
            nbT[idx] = (uint16_t)cc.cell_at(offsets[idx - c * n_off]);
        }
        __syncthreads();
    }
    // no block-wide barrier below this line: every warp is an independent worker
    unsigned char* base = kb_smem + prm.tab_bytes + prm.nbt_bytes + (size_t)warp * prm.rep_bytes;
    uint16_t* p2 = reinterpret_cast<uint16_t*>(base + prm.sm_p2);
    uint8_t* lat = base + prm.sm_lat;
    int32_t* nS = reinterpret_cast<int32_t*>(base + prm.sm_ns);
    double* prodS = reinterpret_cast<double*>(base + prm.sm_prod);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(base + prm.sm_mbar);
    if (prm.use_bulk) {
        if (lane == 0) {
            kb_mbar_init(mbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    uint32_t mbar_phase = 0;
    (void)wpc;

  for (;;) {  // ---- persistent worker loop: one (epoch, replica) item per iteration -------------------------
    int item = 0;
    if (lane == 0) item = atomicAdd(prm.work_counter, 1);
    item = __shfl_sync(KB_FULL, item, 0);
    if (item >= prm.n_items) break;
    const int epoch = item / prm.R;
    const int rep = item - epoch * prm.R;
    long long my_steps = prm.chunk;
    if ((long long)epoch * prm.chunk + my_steps > prm.nsteps) my_steps = prm.nsteps - (long long)epoch * prm.chunk;
    if (epoch > 0) {
        // the previous epoch of this replica was handed out earlier to a resident warp: wait for its write-back
        if (lane == 0) {
            while (kb_ld_acquire(prm.done + rep) < epoch) __nanosleep(128);
        }
        __syncwarp();
        __threadfence();
        kb_fence_proxy_async_all();
    }

    const int P = prm.n_proc, C = prm.ncells, cap = prm.cap, spuck = prm.spuck;
    unsigned char* g_img = reinterpret_cast<unsigned char*>(prm.image) + (size_t)rep * prm.img_bytes;
    // plane 1: in shared memory (staged with the rest of the image) or left in HBM/L2 (P1G): the lists are
    // touched ~10x per step while plane 2 and the lattice take ~40 probes, so keeping only the latter in
    // shared memory trades a few L2 round trips per step for 2-3x more resident replicas per SM
    unsigned char* p1 = P1G ? g_img : base;
    if (P1G) {
        // keep the replica's list base in registers: ptxas otherwise re-derives it (64-bit multiply of the
        // replica index) in every round to stay under the register cap
        unsigned long long p1v = reinterpret_cast<unsigned long long>(p1);
        asm volatile("" : "+l"(p1v));
        p1 = reinterpret_cast<unsigned char*>(p1v);
    }
    unsigned char* p1hi = p1 + prm.off_hi;
    unsigned char* g_stage = g_img + prm.stage_off;
    uint8_t* g_lat = prm.lattice + (size_t)rep * prm.lat_stride;
    int32_t* g_ns = prm.nsites + (size_t)rep * P;

    // ---- stage the replica into shared memory (the image is laid out exactly like p1|p2 here) ---------
    if (prm.use_bulk) {
        if (lane == 0) {
            kb_mbar_expect_tx(mbar, (uint32_t)(prm.stage_bytes + prm.lat_stride));
            kb_bulk_g2s(base, g_stage, (uint32_t)prm.stage_bytes, mbar);
            kb_bulk_g2s(lat, g_lat, (uint32_t)prm.lat_stride, mbar);
        }
        kb_mbar_wait(mbar, mbar_phase);
        mbar_phase ^= 1u;
    } else {
        const uint4* s1 = reinterpret_cast<const uint4*>(g_stage);
        uint4* d1 = reinterpret_cast<uint4*>(base);
        for (int i = lane; i < prm.stage_bytes / 16; i += 32) d1[i] = s1[i];
        const uint4* sl = reinterpret_cast<const uint4*>(g_lat);
        uint4* dl = reinterpret_cast<uint4*>(lat);
        for (int i = lane; i < prm.lat_stride / 16; i += 32) dl[i] = sl[i];
    }
    for (int i = lane; i < P; i += 32) nS[i] = g_ns[i];
    __syncwarp();

    // ---- per-lane process registers -----------------------------------------------------------------
    const int q0 = lane, q1 = lane + 32;
    const bool has0 = (PPL == 2) || q0 < P, has1 = (PPL == 2) && (q1 < P);  // PPL == 2: more than 32 processes
    const double rate0 = has0 ? prm.rates[(size_t)rep * P + q0] : 0.0;
    const double rate1 = has1 ? prm.rates[(size_t)rep * P + q1] : 0.0;
    double integ0 = has0 ? prm.integ[(size_t)rep * P + q0] : 0.0;
    double integ1 = has1 ? prm.integ[(size_t)rep * P + q1] : 0.0;
    const long long ps0 = has0 ? prm.procstat[(size_t)rep * P + q0] : 0;
    const long long ps1 = has1 ? prm.procstat[(size_t)rep * P + q1] : 0;

    KbScalars* const scp = prm.sc + rep;  // only the fields that change are kept in registers / written back
    double kmc_time = scp->kmc_time, kmc_dt = scp->kmc_time_step;
    long long kmc_step = scp->kmc_step;
    int status = scp->status;
    const uint32_t replica_id = scp->replica;
    const uint32_t k0 = (uint32_t)scp->seed, k1 = (uint32_t)(scp->seed >> 32);
    const uint32_t my_off = lane < n_off ? offsets[lane] : 0u;
    // chain lengths padded to a multiple of 2: the extra leading zeros are exact and the unrolled loops need
    // no remainder
    const int len1 = (min(P, 32) + 1) & ~1, len2 = (max(P - 32, 0) + 1) & ~1;
    double* Z1 = prodS;       // [len1 zeros][x_0 .. ]  (64 doubles reserved)
    double* Z2 = prodS + 64;  // [len2 zeros][x_32 ..]  (64 doubles reserved)
    for (int i = lane; i < 128; i += 32) prodS[i] = 0.0;
    __syncwarp();
    const int lastp = P - 1;

    uint32_t cnt0 = 0, cnt1 = 0;      // events of this lane's processes in this item (< 2^32 steps per item)
    double rng_a = 0.0, rng_b = 0.0;  // even lanes: (-log(ran_time), ran_proc); odd lanes: (ran_site, -)

    const long long step0 = kmc_step;  // kmc_step = step0 + it inside the loop (items are < 2^30 steps)
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
            rng_a = (lane & 1) ? u0 : -log(u0);  // odd: ran_site in [0,1); even: -log(ran_time), ran_time in (0,1]
            rng_b = u1;                          // even: ran_proc
        }
        const double neg_log_u = __shfl_sync(KB_FULL, rng_a, 2 * sub);
        const double ran_proc = __shfl_sync(KB_FULL, rng_b, 2 * sub);
        const double ran_site = __shfl_sync(KB_FULL, rng_a, 2 * sub + 1);

        // -- update_accum_rate: products through shared memory, serial float64 recurrence per lane
        const int n0 = has0 ? nS[q0] : 0;
        const int n1 = has1 ? nS[q1] : 0;
        const double pr0 = __dmul_rn((double)n0, rate0);
        const double pr1 = __dmul_rn((double)n1, rate1);
        // Lane L must end up with accum_rates(L+1) = ((x_0 + x_1) + ...) + x_L.  The products sit behind len1
        // leading zeros; lane L starts L+1 entries in, so after len1 additions it has added zeros (exact)
        // followed by x_0..x_L in order -- no per-lane masking, one LDS + one DADD per process.
        // Full first segment (len1 == 32, e.g. RuO2's 36 processes): most products are zero at any time (RuO2
        // sweep: 7 of 36 non-zero on average) and adding 0.0 is exact, so the non-zero products are packed in
        // process order and lane L adds the first popc(nz & lanes <= L) of them: the chain covers the
        // non-zero products only (8, 16 or 32 additions, leading zeros make up the difference).
        const bool full1 = (PPL == 2) || len1 == 32;  // compile-time for the two-segment instantiations
        const unsigned nz0 = full1 ? __ballot_sync(KB_FULL, pr0 != 0.0) : 0u;
        if (full1) {
            if (pr0 != 0.0) Z1[32 + __popc(nz0 & kb_lanemask_lt())] = pr0;
        } else if (has0) {
            Z1[len1 + q0] = pr0;
        }
        if (has1) Z2[len2 + lane] = pr1;
        __syncwarp();
        double acc0 = 0.0;
        {
            if (full1) {
                // chain length in tiers (8, 16, 32: straight-line code, leading zeros make up the difference)
                const int c0 = __popc(nz0);
                const double* top = Z1 + 32 + __popc(nz0 & kb_lanemask_le());
                if (c0 <= 8) {
                    const double* src = top - 8;
#pragma unroll
                    for (int t = 0; t < 8; ++t) acc0 = __dadd_rn(acc0, src[t]);
                } else if (c0 <= 16) {
                    const double* src = top - 16;
#pragma unroll
                    for (int t = 0; t < 16; ++t) acc0 = __dadd_rn(acc0, src[t]);
                } else {
                    const double* src = top - 32;
#pragma unroll
                    for (int t = 0; t < 32; ++t) acc0 = __dadd_rn(acc0, src[t]);
                }
            } else {
                const double* src = Z1 + lane + 1;
                for (int t = 0; t < len1; t += 2) acc0 = __dadd_rn(__dadd_rn(acc0, src[t]), src[t + 1]);
            }
        }
        double acc1 = 0.0;
        if (PPL == 2) {
            acc1 = __shfl_sync(KB_FULL, acc0, 31);  // accum_rates(32)
            const double* src = Z2 + lane + 1;
            if (len2 == 4) {  // 35 or 36 processes (RuO2): straight-line, no generic unrolled loop with its remainders
                acc1 = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(acc1, src[0]), src[1]), src[2]), src[3]);
            } else {
#pragma unroll 1
                for (int t = 0; t < len2; t += 2) acc1 = __dadd_rn(__dadd_rn(acc1, src[t]), src[t + 1]);
            }
        }
        const double total = __shfl_sync(KB_FULL, (PPL == 2 && lastp >= 32) ? acc1 : acc0, lastp & 31);
        if (!(total > 0.0)) { status = KB_DEADLOCK; break; }

        // -- update_clocks / update_integ_rate
        kmc_dt = neg_log_u / total;
        kmc_time = __dadd_rn(kmc_time, kmc_dt);
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
        if (nsel <= 0) { status = KB_DEADLOCK; ++it; break; }  // the clock has advanced: the step counts
        int k = (int)__dadd_rn(1.0, __dmul_rn(ran_site, (double)nsel));
        k = min(k, nsel);
        const uint32_t spi = procinfo[pidx];
        const int cell = kb_p1_get<SPLIT>(p1, p1hi, kb_slot((int)(spi & 63u), (int)((spi >> 6) & 1u), cap, k - 1));
        cnt0 += (pidx == lane);
        if (PPL == 2) cnt1 += (pidx == lane + 32);

        // -- run_proc_nr(pidx+1, site): lattice writes, then rounds of list operations
        const uint4 eh = events[2 * pidx];
        const int ops_start = (int)(eh.x & 0xFFFFu), n_rounds = (int)((eh.x >> 16) & 15u), n_writes = (int)((eh.x >> 20) & 15u);
        int nb;  // lane l: cell index of neighbour offset l
        if (NBT) {
            nb = lane < n_off ? (int)nbT[cell * n_off + lane] : 0;
        } else {
            cc.decode(cell);
            nb = cc.cell_at(my_off);
        }
        {
            const uint32_t w = reinterpret_cast<const uint32_t*>(events + 2 * pidx + 1)[lane & 3];  // lane < 4: its write
            const int wcell = __shfl_sync(KB_FULL, nb, (int)(w & 31u));
            if (lane < n_writes) {
                const int idx = wcell * spuck + (int)((w >> 5) & 7u) - 1;
                const int found = lat[idx], oldsp = (int)((w >> 8) & 15u), newsp = (int)((w >> 12) & 15u);
                if (found != oldsp) {  // replace_species consistency check (base.mpy:1205)
                    status = KB_SPECIES_MISMATCH;
                } else {
                    lat[idx] = (uint8_t)newsp;
                }
            }
        }
        // cumulative op counts of the rounds, one byte each; bytes past the last round repeat the total
        unsigned long long ends = ((unsigned long long)eh.z << 32) | eh.y;
        // Software pipeline over the rounds: the op words of round r+1 (read-only tables) and the cells they
        // name are fetched while round r's list updates are in flight; only nr_of_sites, the class entries
        // and the lists themselves have to wait for the __syncwarp between rounds.
        int endr = (int)((uint32_t)ends & 255u);
        bool valid = lane < endr;
        const uint32_t* op = ops + (ops_start + lane) * STRIDE;
        uint32_t h = valid ? op[0] : 0u, h1 = valid ? op[1] : 0u;
        uint32_t cw[NCOND > 0 ? NCOND : 1];
#pragma unroll
        for (int j = 0; j < NCOND; ++j) cw[j] = (valid && j < (int)((h >> 1) & 7u)) ? op[2 + j] : 0u;
        int ca = __shfl_sync(KB_FULL, nb, (int)__byte_perm(h, 0u, 0x4442u));
        for (int r = 0; r < n_rounds; ++r) {
            // -- fetch round r+1
            ends >>= 8;
            const int end_n = (int)((uint32_t)ends & 255u);
            const bool valid_n = endr + lane < end_n;  // empty after the last round (same total, or 0)
            const uint32_t* op_n = ops + (ops_start + endr + lane) * STRIDE;
            const uint32_t h_n = valid_n ? op_n[0] : 0u, h1_n = valid_n ? op_n[1] : 0u;
            uint32_t cw_n[NCOND > 0 ? NCOND : 1];
#pragma unroll
            for (int j = 0; j < NCOND; ++j) cw_n[j] = valid_n ? op_n[2 + j] : 0u;
            // -- round r
            bool ok = valid;
            const int ncond = (int)((h >> 1) & 7u);
            int32_t* const nSq = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(nS) + __byte_perm(h, 0u, 0x4441u));
            const uint32_t member = h >> 24;
            const bool down = (h >> 4) & 1u;
            const int slot0 = (int)(h1 >> 16);
            uint16_t* entry = p2 + (h1 & 0xFFFFu);
            // this lane is the only one touching list q in this round: its length can be read up front, and
            // when the lists live in L2 the element a del would move is requested before the probes so that
            // the round trip overlaps them (speculative: harmless if the del does not fire)
            const int nq = valid ? *nSq : 0;
            int last = 0;
            if (P1G) {  // unconditional load from an always-valid slot: nothing consumes it before the del body
                const bool want = valid && !(h & 1u) && nq > 0;
                last = kb_p1_get<SPLIT>(p1, p1hi, want ? (down ? slot0 - (nq - 1) : slot0 + (nq - 1)) : 0);
            }
#pragma unroll
            for (int j = 0; j < NCOND; ++j) {
                const int ccell = __shfl_sync(KB_FULL, nb, (int)(cw[j] & 31u));
                if (valid && j < ncond) {
                    const uint32_t sp = lat[ccell * spuck + (int)((cw[j] >> 5) & 7u) - 1];
                    ok = ok && (((cw[j] >> 8) >> sp) & 1u);
                }
            }
            if (P1G) {
                // add_proc (base.mpy:268-302) and the guarded del_proc (base.mpy:211-265) as one predicated
                // sequence when the lists live in L2 (P1G): a round usually mixes both, and the divergent branches cost
                // more than the stores
                const uint32_t e = entry[ca];  // idle lanes read a valid entry (class base 0, lane 0's cell)
                const bool is_add = h & 1u;
                const bool add_bad = ok && is_add && (nq >= C || e != 0);
                const bool add_go = ok && is_add && !add_bad;
                const bool del_go = ok && !is_add && (e >> KB_POS_BITS) == member;
                const int pos = (int)(e & KB_POS_MASK);
                const bool move = del_go && pos < nq;  // the last element takes the freed position
                if (add_bad) status = KB_CAPACITY;
                if (add_go || move) {
                    const int kk = add_go ? nq : pos - 1;
                    kb_p1_set<SPLIT>(p1, p1hi, down ? slot0 - kk : slot0 + kk, add_go ? ca : last);
                }
                if (move) entry[last] = (uint16_t)((member << KB_POS_BITS) | (uint32_t)pos);
                if (add_go || del_go) {
                    entry[ca] = add_go ? (uint16_t)((member << KB_POS_BITS) | (uint32_t)(nq + 1)) : (uint16_t)0;
                    *nSq = add_go ? nq + 1 : nq - 1;
                }
            } else if (ok) {  // lists in shared memory (mini_101: measured 6 % faster with the plain branches)
                if (h & 1u) {  // add_proc (base.mpy:268-302)
                    if (nq >= C || entry[ca] != 0) {
                        status = KB_CAPACITY;
                    } else {
                        kb_p1_set<SPLIT>(p1, p1hi, down ? slot0 - nq : slot0 + nq, ca);
                        entry[ca] = (uint16_t)((member << KB_POS_BITS) | (uint32_t)(nq + 1));
                        *nSq = nq + 1;
                    }
                } else {  // guarded del_proc (base.mpy:211-265)
                    const uint32_t e = entry[ca];
                    if ((e >> KB_POS_BITS) == member) {
                        const int pos = (int)(e & KB_POS_MASK);
                        if (!P1G) last = kb_p1_get<SPLIT>(p1, p1hi, down ? slot0 - (nq - 1) : slot0 + (nq - 1));
                        if (pos < nq) {
                            kb_p1_set<SPLIT>(p1, p1hi, down ? slot0 - (pos - 1) : slot0 + (pos - 1), last);
                            entry[last] = (uint16_t)((member << KB_POS_BITS) | (uint32_t)pos);
                        }
                        entry[ca] = 0;
                        *nSq = nq - 1;
                    }
                }
            }
            // -- rotate
            ca = __shfl_sync(KB_FULL, nb, (int)__byte_perm(h_n, 0u, 0x4442u));
            h = h_n; h1 = h1_n; valid = valid_n; endr = end_n;
#pragma unroll
            for (int j = 0; j < NCOND; ++j) cw[j] = cw_n[j];
            __syncwarp();
        }
        __syncwarp();  // lattice writes of an event without ops must be visible to the next step
        // a lane-local failure (species mismatch / capacity) stops the replica for every lane
        if (__any_sync(KB_FULL, status != KB_OK)) {
            const unsigned mm = __ballot_sync(KB_FULL, status == KB_SPECIES_MISMATCH);
            if (mm) {
                // error tuple of the first failing replace_species call (old, new, found, site, step), rebuilt
                // here so that the step loop carries no registers for it; the lattice site was left untouched
                const int src = __ffs(mm) - 1;
                const uint32_t w = reinterpret_cast<const uint32_t*>(events + 2 * pidx + 1)[src];
                const int wcell = __shfl_sync(KB_FULL, nb, (int)(w & 31u));
                if (lane == 0) {
                    const int idx = wcell * spuck + (int)((w >> 5) & 7u) - 1;
                    scp->err[0] = (int)((w >> 8) & 15u); scp->err[1] = (int)((w >> 12) & 15u);
                    scp->err[2] = lat[idx]; scp->err[3] = idx + 1; scp->err[4] = (int)(step0 + it + 1);
                }
            }
            status = __reduce_max_sync(KB_FULL, status);
        }
    }
    status = __reduce_max_sync(KB_FULL, status);
    kmc_step = step0 + it;

    // ---- write back -----------------------------------------------------------------------------------
    __syncwarp();
    if (prm.use_bulk) {
        kb_fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            kb_bulk_s2g(g_stage, base, (uint32_t)prm.stage_bytes);
            kb_bulk_s2g(g_lat, lat, (uint32_t)prm.lat_stride);
            kb_bulk_commit_wait();
            kb_fence_proxy_async_all();
        }
    } else {
        uint4* s1 = reinterpret_cast<uint4*>(g_stage);
        const uint4* d1 = reinterpret_cast<const uint4*>(base);
        for (int i = lane; i < prm.stage_bytes / 16; i += 32) s1[i] = d1[i];
        uint4* sl = reinterpret_cast<uint4*>(g_lat);
        const uint4* dl = reinterpret_cast<const uint4*>(lat);
        for (int i = lane; i < prm.lat_stride / 16; i += 32) sl[i] = dl[i];
    }
    for (int i = lane; i < P; i += 32) g_ns[i] = nS[i];
    if (has0) { prm.integ[(size_t)rep * P + q0] = integ0; prm.procstat[(size_t)rep * P + q0] = ps0 + cnt0; }
    if (has1) { prm.integ[(size_t)rep * P + q1] = integ1; prm.procstat[(size_t)rep * P + q1] = ps1 + cnt1; }
    if (lane == 0) {
        scp->kmc_time = kmc_time; scp->kmc_time_step = kmc_dt; scp->kmc_step = kmc_step; scp->status = status;
    }
    // publish: everything this warp wrote for the replica is visible before the epoch counter moves
    __threadfence();
    __syncwarp();
    if (lane == 0) kb_st_release(prm.done + rep, epoch + 1);
    __syncwarp();  // the shared-memory block is reused by the next item
  }
}
