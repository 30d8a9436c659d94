// kmos_b200.cu -- C-ABI (include/kmos_b200.h) over the two CUDA step engines.
//
//   kb_generic_kernel   thread-per-replica byte-code engine on HBM-resident state (kb_interp.h):
//                       initialize_state / touchup, lat_int, otf, lattices that do not fit shared memory
//   kb_smem_kernel      warp-per-replica shared-memory engine (kb_smem.cuh): local_smart models that fit
//
// Device state per batch (R replicas), all in HBM between launches:
//   lattice  uint8  [R][lat_stride]          species per site, site-number order
//   p1, p2   idx_t  [R][plane]               avail_sites planes per process and cell (kb_common.h)
//   nsites   int32  [R][P]    rates/integ/accum double [R][P]    procstat int64 [R][P]
//   scalars  KbScalars[R]     kmc_time, kmc_time_step, kmc_step, Philox key/replica id, status, error tuple
//   otf only: rates_matrix double [R][P][ncells+1], accum_proc double [R][ncells], lut double [R][lut]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/kmos_b200.h"
#include "kb_interp.h"
#include "kb_smem.cuh"
#include "kb_latint.cuh"
#include "kb_otf.cuh"
#include "kb_gen.cuh"
#include "kb_otf_fast.cuh"

#include <dlfcn.h>

static thread_local std::string g_err;
static int set_err(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return set_err(KMOS_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

struct kmos_b200_batch;
static cudaError_t kb_h2d(kmos_b200_batch* b, void* dst, const void* src, size_t bytes);

struct kmos_b200_model {
    std::vector<int32_t> blob;
    KbModelView h;  // pointers into blob (host)
    bool dev_supported;
    int max_off[3];  // largest |offset| per axis in the device tables
};

typedef void (*kb_smem_fn)(const KbSmemParams);

// an exporter-generated proclist module (kmos_b200/codegen.py -> proclist_<model>_<hash>.so), see kb_gen.cuh
struct KbGenModule {
    void* handle = nullptr;
    const KbGenInfo* info = nullptr;
    int (*plan)(const int32_t*, int32_t, int32_t, KbGenPlan*) = nullptr;
    int (*build_tables)(const KbGenPlan*, uint32_t*) = nullptr;
    int (*launch)(const KbGenParams*, const KbGenPlan*, int*, int, int, void*) = nullptr;
};

struct kmos_b200_batch {
    kmos_b200_model* model;
    int R, device;
    KbGeom g;
    bool idx32;
    int lat_stride;
    size_t plane_bytes;  // bytes of one avail plane per replica (multiple of 16)
    cudaStream_t stream, own_stream;
    cudaEvent_t ev0, ev1;
    int32_t* d_blob;
    KbModelView d;  // pointers into d_blob
    uint8_t* lattice;
    void *p1, *p2;
    int32_t* nsites;
    double *rates, *integ, *accum;
    int64_t* procstat;
    KbScalars* sc;
    double *rates_matrix, *accum_proc, *lut;
    uint16_t* image;     // compact avail planes of the shared-memory engine [R][img_bytes/2]
    bool compact_valid;  // true: `image` holds the avail tables, p1/p2 are stale; false: the other way round
    double* tally;
    double* occ;  // scratch of reduce_tallies: occupation[R][n_species*spuck]
    int32_t* group_of;
    int tally_groups;
    // kernel choice
    int kernel;  // KMOS_B200_KERNEL_GENERIC / SMEM
    KbSmemParams sp;
    int wpc, ctas_per_sm, sm_count, smem_bytes, ppl, ncond, regs;
    kb_smem_fn fn;
    std::vector<int32_t> spec;  // lane tables specialised for this geometry (host copy)
    int32_t* d_spec;
    int* d_sched;  // [0] work counter, [1..R] finished epochs per replica (persistent scheduling)
    bool smem_ok;
    double smem_score;  // resident replicas per SM, discounted by the placement's latency penalty
    std::string smem_reason;
    bool otf_ok; // warp-per-replica otf kernel available (lane per rates_matrix row)
    bool li_ok;  // warp-per-replica lat_int kernel available
    KbLatintParams li;
    int li_wpc, li_smem_bytes, li_mode;  // li_mode 0: lat_int decision trees, 1: local_smart flattened ops
    int li_lat_bytes;                    // > 0: bytes per warp of the 4-bit lattice copy in shared memory (else 0)
    // generated per-model kernel (kmos_b200_batch_attach_proclist)
    KbGenModule gen;
    bool gen_ok;
    bool initialised;          // init_state / set_configuration / reload_replica has run
    KbGenPlan gp;
    uint32_t* d_gen_tab;
    uint32_t* d_gen_writes;
    int32_t* d_gen_dev;        // header + procinfo words for the pack/unpack kernels on the generated kernel's image
    unsigned char* gen_image;  // [R][gp.img_bytes]
    int compact_kind;          // which compact image is valid when compact_valid: 0 interpreter, 1 generated
};

extern "C" const char* kmos_b200_last_error(void) { return g_err.c_str(); }

extern "C" int kmos_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---------------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------------
extern "C" int kmos_b200_model_create(const int32_t* blob, int64_t n_words, kmos_b200_model** out) {
    if (!blob || !out || n_words < 14) return set_err(KMOS_B200_ERR_ARG, "model_create: bad arguments");
    kmos_b200_model* m = new kmos_b200_model;
    m->blob.assign(blob, blob + n_words);
    if (!kb_model_view(m->blob.data(), n_words, m->blob.data(), &m->h)) {
        delete m;
        return set_err(KMOS_B200_ERR_MODEL, "model_create: not a KB20 v4 table image");
    }
    for (int q = 0; q < m->h.n_proc; ++q)
        if (m->h.procsite[q] < 1 || m->h.procsite[q] > m->h.spuck) {
            delete m;
            return set_err(KMOS_B200_ERR_UNSUPPORTED,
                           "model_create: a process is registered on several site types (per-cell avail layout)");
        }
    if (m->h.n_species > 32) { delete m; return set_err(KMOS_B200_ERR_UNSUPPORTED, "more than 32 species"); }
    m->dev_supported = m->h.dev && m->h.dev_len >= 16 && m->h.dev[0] == 2 && m->h.dev[1] == 1;
    m->max_off[0] = m->max_off[1] = m->max_off[2] = 0;
    if (m->dev_supported) {
        const int32_t* d = m->h.dev;
        for (int i = 0; i < d[8]; ++i) {
            uint32_t w = (uint32_t)d[d[7] + i];
            for (int a = 0; a < 3; ++a) {
                int v = abs((int)(int8_t)((w >> (8 * a)) & 255u));
                if (v > m->max_off[a]) m->max_off[a] = v;
            }
        }
    }
    *out = m;
    return KMOS_B200_OK;
}
extern "C" void kmos_b200_model_destroy(kmos_b200_model* m) { delete m; }
extern "C" int kmos_b200_model_nproc(const kmos_b200_model* m) { return m->h.n_proc; }
extern "C" int kmos_b200_model_nspecies(const kmos_b200_model* m) { return m->h.n_species; }
extern "C" int kmos_b200_model_spuck(const kmos_b200_model* m) { return m->h.spuck; }
extern "C" int kmos_b200_model_lut_size(const kmos_b200_model* m) { return m->h.lut_total; }

// ---------------------------------------------------------------------------------------------------
// generic kernels (thread per replica)
// ---------------------------------------------------------------------------------------------------
struct KbBatchView {
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
};

template <typename idx_t>
__device__ __forceinline__ void kb_load_replica(const KbBatchView& b, int rep, KbReplica<idx_t>& r) {
    const int P = b.m.n_proc;
    r.lattice = b.lattice + (size_t)rep * b.lat_stride;
    r.nsites = b.nsites + (size_t)rep * P;
    r.p1 = reinterpret_cast<idx_t*>(b.p1) + (size_t)rep * b.plane_elems;
    r.p2 = reinterpret_cast<idx_t*>(b.p2) + (size_t)rep * b.plane_elems;
    r.rates = b.rates + (size_t)rep * P;
    r.integ = b.integ + (size_t)rep * P;
    r.accum = b.accum + (size_t)rep * P;
    r.procstat = b.procstat + (size_t)rep * P;
    r.rates_matrix = b.rates_matrix ? b.rates_matrix + (size_t)rep * P * (b.g.ncells + 1) : nullptr;
    r.accum_proc = b.accum_proc ? b.accum_proc + (size_t)rep * b.g.ncells : nullptr;
    r.lut = b.lut ? b.lut + (size_t)rep * (b.m.lut_total > 0 ? b.m.lut_total : 1) : nullptr;
    const KbScalars s = b.sc[rep];
    r.kmc_time = s.kmc_time; r.kmc_time_step = s.kmc_time_step; r.kmc_step = s.kmc_step;
    r.seed = s.seed; r.replica = s.replica; r.status = s.status;
    for (int i = 0; i < 5; ++i) r.err[i] = s.err[i];
}
template <typename idx_t>
__device__ __forceinline__ void kb_store_replica(const KbBatchView& b, int rep, const KbReplica<idx_t>& r) {
    KbScalars s = b.sc[rep];
    s.kmc_time = r.kmc_time; s.kmc_time_step = r.kmc_time_step; s.kmc_step = r.kmc_step; s.status = r.status;
    for (int i = 0; i < 5; ++i) s.err[i] = r.err[i];
    b.sc[rep] = s;
}

enum { KB_MODE_STEPS = 0, KB_MODE_INIT = 1, KB_MODE_ADJUST = 2, KB_MODE_ACCUM = 3, KB_MODE_NEXT = 4, KB_MODE_RUNPROC = 5,
       KB_MODE_RECALC = 6 };

template <typename idx_t>
__global__ void kb_generic_kernel(const KbBatchView b, int mode, long long n, int layer, int only_rep,
                                  int32_t* io_proc, int32_t* io_site) {
    int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= b.R) return;
    if (only_rep >= 0 && rep != only_rep) return;
    KbReplica<idx_t> r;
    kb_load_replica(b, rep, r);
    KbInterp<idx_t> it(b.m, b.g, r);
    if (mode == KB_MODE_STEPS) it.do_kmc_steps(n);
    else if (mode == KB_MODE_INIT) it.init_state(layer);
    else if (mode == KB_MODE_ADJUST) it.adjust_database(layer);
    else if (mode == KB_MODE_RECALC) it.recalculate_rates_matrix();
    else if (mode == KB_MODE_NEXT) {
        // get_next_kmc_step (proclist_generic_subroutines.mpy:85-110): the step's uniforms, no clock update,
        // and -- as in the reference -- ran_time is what selects the site
        double ran_time, ran_proc, ran_site;
        kb_philox_step(r.seed, r.replica, (uint64_t)r.kmc_step, &ran_time, &ran_proc, &ran_site);
        it.update_accum_rate();
        int proc = 0, cell = 0;
        io_proc[rep] = 0; io_site[rep] = 0;
        if (r.status == KB_OK && r.accum[b.m.n_proc - 1] > 0. && it.determine_procsite(ran_proc, ran_time, &proc, &cell)) {
            io_proc[rep] = proc;
            io_site[rep] = b.m.spuck * cell + it.anchor_n(proc);
        } else if (r.status == KB_OK) {
            it.fail(KB_DEADLOCK);
        }
    } else if (mode == KB_MODE_RUNPROC) {
        // run_proc_nr(proc, nr_site): procstat + the event; no random numbers, no clock
        const int proc = io_proc[rep], site = io_site[rep];
        if (proc >= 1 && proc <= b.m.n_proc && site >= 1 && site <= b.g.volume && r.status == KB_OK)
            it.run_proc_nr(proc, (site - 1) / b.m.spuck);
    } else it.update_accum_rate();
    kb_store_replica(b, rep, r);
}

// get_occupation (proclist_generic_subroutines.mpy:113-158): counts[R][n_species][spuck] / ncells
__global__ void kb_occupation_kernel(const uint8_t* lattice, int lat_stride, int R, int volume, int spuck,
                                     int n_species, int ncells, double* out) {
    int rep = blockIdx.x;
    extern __shared__ int kb_occ[];
    for (int i = threadIdx.x; i < n_species * spuck; i += blockDim.x) kb_occ[i] = 0;
    __syncthreads();
    const uint8_t* lat = lattice + (size_t)rep * lat_stride;
    for (int i = threadIdx.x; i < volume; i += blockDim.x) {
        int s = lat[i];
        if (s < n_species) atomicAdd(&kb_occ[s * spuck + (i % spuck)], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_species * spuck; i += blockDim.x)
        out[(size_t)rep * n_species * spuck + i] = (double)kb_occ[i] / (double)ncells;
}

// tallies per group: [P] procstat | [P] integ | [ns*spuck] occupation | kmc_time | kmc_steps | n_replicas
// One block per group; thread w owns word w and adds the group's replicas in replica order (deterministic).
// The group's members are first gathered, tile by tile and in replica order, into shared memory (ballot + rank
// over the block), so a thread walks its group's replicas only, not all R group ids: 16 384 replicas in 256
// groups took 0.87 ms per launch (2.7 % of a bench step) with every thread scanning group_of[] itself.
#define KB_TALLY_THREADS 128
#define KB_TALLY_TILE 2048
__global__ void __launch_bounds__(KB_TALLY_THREADS) kb_tally_kernel(const KbScalars* sc, const int64_t* procstat, const double* integ,
                                                                    const double* occ, const int32_t* group_of, int R, int P, int nocc,
                                                                    double* out, int words) {
    __shared__ int member[KB_TALLY_TILE];
    __shared__ int wcount[KB_TALLY_THREADS / 32];
    const int grp = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w0 = 0; w0 < words; w0 += KB_TALLY_THREADS) {  // uniform trip count: the block synchronises inside
        const int w = w0 + threadIdx.x;
        double acc = 0.0;
        for (int t0 = 0; t0 < R; t0 += KB_TALLY_TILE) {
            const int t1 = min(R, t0 + KB_TALLY_TILE);
            int n_mem = 0;
            if (group_of) {
                for (int base = t0; base < t1; base += KB_TALLY_THREADS) {
                    const int rep = base + threadIdx.x;
                    const bool m = rep < t1 && group_of[rep] == grp;
                    const unsigned bal = __ballot_sync(0xffffffffu, m);
                    if (lane == 0) wcount[warp] = __popc(bal);
                    __syncthreads();
                    int before = 0, all = 0;
                    for (int k = 0; k < KB_TALLY_THREADS / 32; ++k) {
                        before += k < warp ? wcount[k] : 0;
                        all += wcount[k];
                    }
                    if (m) member[n_mem + before + __popc(bal & ((1u << lane) - 1u))] = rep;
                    n_mem += all;
                    __syncthreads();
                }
            } else {
                n_mem = grp == 0 ? t1 - t0 : 0;  // no group ids: every replica belongs to group 0, no list needed
            }
            if (w < words) {
                for (int i = 0; i < n_mem; ++i) {
                    const int rep = group_of ? member[i] : t0 + i;
                    double v;
                    if (w < P) v = (double)procstat[(size_t)rep * P + w];
                    else if (w < 2 * P) v = integ[(size_t)rep * P + (w - P)];
                    else if (w < 2 * P + nocc) v = occ[(size_t)rep * nocc + (w - 2 * P)];
                    else if (w == 2 * P + nocc) v = sc[rep].kmc_time;
                    else if (w == 2 * P + nocc + 1) v = (double)sc[rep].kmc_step;
                    else v = 1.0;
                    acc += v;
                }
            }
            __syncthreads();  // the list is rewritten by the next tile
        }
        if (w < words) out[(size_t)grp * words + w] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------
// batch
// ---------------------------------------------------------------------------------------------------
static KbBatchView batch_view(const kmos_b200_batch* b) {
    KbBatchView v;
    v.m = b->d; v.g = b->g; v.R = b->R; v.lat_stride = b->lat_stride;
    v.plane_elems = b->plane_bytes / (b->idx32 ? 4 : 2);
    v.lattice = b->lattice; v.p1 = b->p1; v.p2 = b->p2; v.nsites = b->nsites; v.rates = b->rates;
    v.integ = b->integ; v.accum = b->accum; v.procstat = b->procstat; v.sc = b->sc;
    v.rates_matrix = b->rates_matrix; v.accum_proc = b->accum_proc; v.lut = b->lut;
    return v;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static bool magic_ok(uint32_t magic, int d, int limit) {
    for (int c = 0; c < limit; ++c)
        if ((int)(((uint64_t)(uint32_t)c * magic) >> 32) != c / d) return false;
    return true;
}

template <int PPL, int NC, bool NBT>
static kb_smem_fn kb_pick3(bool split, bool p1g) {
    if (split) return p1g ? kb_smem_kernel<PPL, NC, true, true, NBT> : kb_smem_kernel<PPL, NC, true, false, NBT>;
    return p1g ? kb_smem_kernel<PPL, NC, false, true, NBT> : kb_smem_kernel<PPL, NC, false, false, NBT>;
}
template <int PPL, int NC>
static kb_smem_fn kb_pick2(bool split, bool p1g, bool nbt) {
    return nbt ? kb_pick3<PPL, NC, true>(split, p1g) : kb_pick3<PPL, NC, false>(split, p1g);
}
template <int PPL>
static kb_smem_fn kb_pick1(int nc, bool split, bool p1g, bool nbt) {
    switch (nc) {
    case 0: return kb_pick2<PPL, 0>(split, p1g, nbt);
    case 1: return kb_pick2<PPL, 1>(split, p1g, nbt);
    case 2: return kb_pick2<PPL, 2>(split, p1g, nbt);
    case 3: return kb_pick2<PPL, 3>(split, p1g, nbt);
    default: return kb_pick2<PPL, 4>(split, p1g, nbt);
    }
}
static kb_smem_fn kb_pick(int ppl, int nc, bool split, bool p1g, bool nbt) {
    return ppl == 2 ? kb_pick1<2>(nc, split, p1g, nbt) : kb_pick1<1>(nc, split, p1g, nbt);
}

// Lane tables specialised for one geometry: every op gets a second word with its class base (cls * ncells)
// and the first slot of its list (arena * cap, or the arena's last slot if the list grows downwards), so the
// kernel's round body needs no multiplications.  Layout otherwise as written by kmos_b200/devtables.py.
static bool specialise_tables(const int32_t* d, int ncells, int cap, std::vector<int32_t>* out) {
    const int n_events = d[2], events_off = d[3], ops_off = d[4], n_ops = d[5], stride = d[6];
    const int offsets_off = d[7], n_off = d[8], procinfo_off = d[9], n_proc = d[2];
    const int ev_words = offsets_off - 0;  (void)ev_words;
    const int new_stride = stride + 1;
    std::vector<int32_t>& o = *out;
    o.assign(d, d + ops_off);  // header + events
    const int new_ops_off = (int)o.size();
    for (int i = 0; i < n_ops; ++i) {
        const uint32_t h = (uint32_t)d[ops_off + i * stride];
        const uint32_t kind = h & 1u, ncond = (h >> 1) & 7u, off_id = (h >> 4) & 31u, q = (h >> 9) & 63u;
        const uint32_t cls = (h >> 15) & 31u, member = (h >> 20) & 7u, arena = (h >> 23) & 63u, dir = (h >> 29) & 1u;
        const uint32_t p2base = cls * (uint32_t)ncells;
        const uint32_t slot0 = arena * (uint32_t)cap + (dir ? (uint32_t)cap - 1u : 0u);
        if (p2base > 0xFFFFu || slot0 > 0xFFFFu) return false;
        o.push_back((int32_t)(kind | (ncond << 1) | (dir << 4) | ((4u * q) << 8) | (off_id << 16) | (member << 24)));
        o.push_back((int32_t)(p2base | (slot0 << 16)));
        for (int j = 1; j < stride; ++j) o.push_back(d[ops_off + i * stride + j]);
    }
    const int new_offsets_off = (int)o.size();
    for (int i = 0; i < n_off; ++i) o.push_back(d[offsets_off + i]);
    const int new_procinfo_off = (int)o.size();
    for (int i = 0; i < n_proc; ++i) o.push_back(d[procinfo_off + i]);
    o[4] = new_ops_off; o[6] = new_stride; o[7] = new_offsets_off; o[9] = new_procinfo_off;
    (void)n_events; (void)events_off;
    return true;
}

// lat_int: warp-per-replica kernel on the canonical HBM layout
static void plan_latint(kmos_b200_batch* b) {
    const kmos_b200_model* m = b->model;
    b->li_ok = false;
    b->otf_ok = false;
    if (m->h.backend == KB_BACKEND_OTF) {
        const int nchunk = (b->g.ncells + KB_OTF_CHUNK - 1) / KB_OTF_CHUNK;
        cudaDeviceProp prop;
        if (m->h.n_proc <= 64 && (long long)m->h.n_proc * nchunk <= b->g.ncells &&
            cudaGetDeviceProperties(&prop, b->device) == cudaSuccess) {
            b->sm_count = prop.multiProcessorCount;
            b->otf_ok = true;
        }
        return;
    }
    const int32_t* d = nullptr;
    int dlen = 0;
    if (m->h.backend == KB_BACKEND_LAT_INT && m->h.dev && m->h.dev_len >= 16 && m->h.dev[0] == 3 && m->h.dev[1] == 1) {
        d = m->h.dev; dlen = m->h.dev_len; b->li_mode = 0;
    } else if (m->h.backend == KB_BACKEND_LOCAL_SMART && m->h.dev_hbm && m->h.dev_hbm_len >= 16 &&
               m->h.dev_hbm[0] == 4 && m->h.dev_hbm[1] == 1) {
        d = m->h.dev_hbm; dlen = m->h.dev_hbm_len; b->li_mode = 1;
    }
    if (!d) return;
    if (m->h.n_proc > 256) return;
    const int li_ppl = m->h.n_proc <= 32 ? 1 : (m->h.n_proc <= 64 ? 2 : (m->h.n_proc <= 128 ? 4 : 8));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, b->device) != cudaSuccess) return;
    // distinct offsets of one event must be distinct cells (the generated code assumes it too: folded probes,
    // one list operation per (process, cell)): no aliasing under periodic wrap
    int max_off[3] = {0, 0, 0};
    for (int i = 0; i < d[8]; ++i) {
        uint32_t w = (uint32_t)d[d[7] + i];
        for (int a = 0; a < m->h.dim; ++a) {
            const int o = abs((int)(int8_t)((w >> (8 * a)) & 255u));
            if (2 * o >= b->g.size[a]) return;
            if (o > max_off[a]) max_off[a] = o;
        }
    }
    int Lx = b->g.size[0], LxLy = b->g.size[0] * b->g.size[1];
    if (Lx == 1 || LxLy == 1) return;
    KbLatintParams& li = b->li;
    memset(&li, 0, sizeof li);
    li.magic_x = (uint32_t)((0x100000000ull / (uint64_t)Lx) + 1);
    li.magic_xy = (uint32_t)((0x100000000ull / (uint64_t)LxLy) + 1);
    if (!magic_ok(li.magic_x, Lx, b->g.ncells) || !magic_ok(li.magic_xy, LxLy, b->g.ncells)) return;
    li.dev_words = dlen;
    li.tab_bytes = (int)align_up((size_t)li.dev_words * 4, 128);
    li.rep_bytes = 768 * li_ppl;  // per 32 processes: nr_of_sites, event counters (128 B each), product buffer (512 B)
    b->li_wpc = 8;
    b->li_smem_bytes = li.tab_bytes + b->li_wpc * li.rep_bytes;
    if (li_ppl >= 4) {  // the kernels for more than 64 processes read their tables in place (kb_latint.cuh)
        li.tab_bytes = 0;
        b->li_smem_bytes = b->li_wpc * li.rep_bytes;
    }
    // lattice copy in shared memory (4 bits per site): the probes of an event are chains of dependent lattice
    // reads.  Taken when as many CTAs stay resident as the batch can use (at most 3, what the registers allow).
    b->li_lat_bytes = 0;
    if (li_ppl <= 2 && m->h.n_species <= 15 && !getenv("KMOS_B200_NO_LATS")) {
        const int lat_bytes = (int)align_up((size_t)((b->g.volume + 7) / 8) * 4, 16);
        const int with_lat = li.tab_bytes + b->li_wpc * (li.rep_bytes + lat_bytes);
        int needed = (b->R + b->li_wpc * prop.multiProcessorCount - 1) / (b->li_wpc * prop.multiProcessorCount);
        if (needed > 3) needed = 3;
        if (with_lat <= (int)prop.sharedMemPerBlockOptin &&
            (int)(prop.sharedMemPerMultiprocessor / (size_t)(with_lat + 1024)) >= needed) {
            b->li_lat_bytes = lat_bytes;
            li.rep_bytes += lat_bytes;
            b->li_smem_bytes = with_lat;
        }
    }
    if (b->li_smem_bytes > (int)prop.sharedMemPerBlockOptin) return;
    li.n_proc = m->h.n_proc; li.n_species = m->h.n_species; li.spuck = m->h.spuck; li.dim = m->h.dim;
    // an event's probes reach two offsets from its cell (an op's anchor plus that op's own probes)
    for (int a = 0; a < 3; ++a) li.margin[a] = a < m->h.dim ? 2 * max_off[a] : 0;
    {   // worth it for the decision-tree walks of lat_int on lattices that are mostly interior (128x128: 94 %);
        // the local_smart instantiations compile without it (kb_latint.cuh), small lattices switch it off here
        double frac = 1.0;
        for (int a = 0; a < m->h.dim; ++a) frac *= (double)std::max(0, b->g.size[a] - 2 * li.margin[a]) / b->g.size[a];
        const char* ienv = getenv("KMOS_B200_INTERIOR");  // 1: whenever a cell qualifies, 0: never (parity tests)
        const bool on = ienv ? atoi(ienv) != 0 : frac >= 0.8;
        if (b->li_mode != 0 || !on)
            for (int a = 0; a < 3; ++a) li.margin[a] = 1 << 20;  // no cell counts as interior
    }
    for (int a = 0; a < 3; ++a) li.size[a] = b->g.size[a];
    li.ncells = b->g.ncells;
    b->sm_count = prop.multiProcessorCount;
    b->li_ok = true;
}

// choose the shared-memory configuration; sets b->smem_ok / smem_reason
static void plan_smem(kmos_b200_batch* b) {
    const kmos_b200_model* m = b->model;
    b->smem_ok = false;
    if (m->h.backend != KB_BACKEND_LOCAL_SMART || !m->dev_supported) { b->smem_reason = "no device tables for this model/backend"; return; }
    if (b->g.ncells > (int)KB_POS_MASK) { b->smem_reason = "more than 8191 cells"; return; }
    if (m->h.n_proc > 64) { b->smem_reason = "more than 64 processes"; return; }
    // an anchor cell must identify its neighbour offset uniquely: no two offsets may alias under wrap
    for (int a = 0; a < m->h.dim; ++a)
        if (2 * m->max_off[a] >= b->g.size[a]) { b->smem_reason = "lattice smaller than twice the interaction range"; return; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, b->device) != cudaSuccess) { b->smem_reason = "no device properties"; return; }
    b->sm_count = prop.multiProcessorCount;
    const int max_smem = (int)prop.sharedMemPerBlockOptin;
    const int per_sm = (int)prop.sharedMemPerMultiprocessor;
    const int32_t* d = m->h.dev;
    KbSmemParams& sp = b->sp;
    memset(&sp, 0, sizeof sp);
    sp.dev_words = m->h.dev_len + d[5];  // specialised ops carry one extra word each
    sp.tab_bytes = (int)align_up((size_t)sp.dev_words * 4, 128);
    // per-CTA neighbour table (cell x offset -> cell), if it is small next to the replicas it serves
    const size_t nbt = (size_t)b->g.ncells * d[8] * 2;
    const char* nbt_env = getenv("KMOS_B200_NO_NBT");
    sp.nbt_bytes = (nbt <= 16384 && !(nbt_env && nbt_env[0] == '1')) ? (int)align_up(nbt, 128) : 0;
    sp.n_classes = d[10]; sp.n_arenas = d[11];
    sp.plane_bytes = (int)b->plane_bytes;
    sp.lat_stride = b->lat_stride;
    b->ppl = m->h.n_proc > 32 ? 2 : 1;
    b->ncond = d[14];
    if (b->ncond > 4) { b->smem_reason = "more than 4 probes per add"; return; }
    // Two placements of plane 1 (the lists): with the rest of the image in shared memory, or left in HBM/L2.
    // spare slots: the two lists of an arena may overshoot ncells transiently by at most the adds of one
    // event (devtables.py).  Split storage (shared-memory placement, <= 512 cells) additionally keeps the
    // lists >= 8 slots apart so that no byte of the bit-8 bitmap is shared between two lists.
    const char* p1env = getenv("KMOS_B200_P1");
    const char* ns_env = getenv("KMOS_B200_NO_SPLIT");
    int best_total = 0;
    double best_score = 0;
    KbSmemParams best = sp;
    for (int p1g = 0; p1g <= 1; ++p1g) {
        if (p1env && ((p1env[0] == 's' && p1g == 1) || (p1env[0] == 'g' && p1g == 0))) continue;
        KbSmemParams c = sp;
        c.p1_global = p1g;
        c.split = (!p1g && b->g.ncells <= 512 && !(ns_env && ns_env[0] == '1')) ? 1 : 0;
        c.cap = c.split ? (int)align_up((size_t)b->g.ncells + d[15] + 8, 8) : (b->g.ncells + d[15] + 1) & ~1;
        const size_t p1_bytes = c.split ? (size_t)c.n_arenas * c.cap : (size_t)c.n_arenas * c.cap * 2;
        c.off_hi = (int)p1_bytes;
        c.off_p2 = (int)align_up(p1_bytes + (c.split ? (size_t)c.n_arenas * c.cap / 8 : 0), 16);
        c.img_bytes = (int)align_up((size_t)c.off_p2 + (size_t)c.n_classes * b->g.ncells * 2, 16);
        c.stage_off = p1g ? c.off_p2 : 0;
        c.stage_bytes = c.img_bytes - c.stage_off;
        c.sm_p2 = p1g ? 0 : c.off_p2;
        c.sm_lat = c.stage_bytes;
        c.sm_ns = c.sm_lat + c.lat_stride;
        c.sm_prod = (int)align_up((size_t)c.sm_ns + 4 * m->h.n_proc, 16);
        c.sm_mbar = c.sm_prod + 8 * 128;  // two zero-prefixed product buffers (kb_smem.cuh)
        c.rep_bytes = (int)align_up((size_t)c.sm_mbar + 16, 128);
        cudaFuncAttributes fa;
        kb_smem_fn fn = kb_pick(b->ppl, b->ncond, c.split != 0, p1g != 0, c.nbt_bytes != 0);
        if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess) { cudaGetLastError(); continue; }
        // registers are allocated per warp in units of 8 per thread and per CTA in groups of 4 warps
        const int regs = (fa.numRegs + 7) & ~7;
        const int reg_warps = prop.regsPerMultiprocessor / (32 * regs);  // warps an SM can hold
        const int max_threads = fa.maxThreadsPerBlock;
        for (int w = 1; w <= 32 && w * 32 <= max_threads; ++w) {
            const int w4 = (w + 3) & ~3;
            if (w4 > reg_warps) break;
            int bytes = c.tab_bytes + c.nbt_bytes + w * c.rep_bytes;
            if (bytes > max_smem) break;
            int n = per_sm / (bytes + 1024);  // 1 KB/CTA reserved by the driver
            if (n > 32) n = 32;
            if (n * w > 64) n = 64 / w;
            if (n * w4 > reg_warps) n = reg_warps / w4;
            if (n < 1) continue;
            // score: resident replicas per SM x how full the last wave of this batch is; the lists-in-L2
            // placement pays ~1.5x more latency per step, so it must bring that many more replicas
            // score: resident replicas per SM (work is handed out dynamically in epochs, so the fill of the
            // last wave does not matter); the lists-in-L2 placement pays ~1.5x more latency per step
            const double score = n * w / (p1g ? 1.5 : 1.0);
            if (score > best_score * 1.0001 || (score > best_score * 0.9999 && n * w < best_total)) {
                best_score = score; best_total = n * w; best = c;
                b->wpc = w; b->ctas_per_sm = n; b->smem_bytes = c.tab_bytes + c.nbt_bytes + w * c.rep_bytes;
                b->fn = fn; b->regs = fa.numRegs;
            }
        }
    }
    if (!best_total) { b->smem_reason = "one replica does not fit in shared memory"; return; }
    sp = best;
    b->smem_score = best_score;
    const char* wenv = getenv("KMOS_B200_WARPS_PER_CTA");
    if (wenv && atoi(wenv) > 0 && sp.tab_bytes + sp.nbt_bytes + atoi(wenv) * sp.rep_bytes <= max_smem) {
        b->wpc = atoi(wenv);
        b->ctas_per_sm = per_sm / (sp.tab_bytes + sp.nbt_bytes + b->wpc * sp.rep_bytes + 1024);
        if (b->ctas_per_sm < 1) b->ctas_per_sm = 1;
        b->smem_bytes = sp.tab_bytes + sp.nbt_bytes + b->wpc * sp.rep_bytes;
    }
    if (!specialise_tables(d, b->g.ncells, sp.cap, &b->spec)) { b->smem_reason = "class/arena bases exceed 16 bits"; return; }
    if ((int)b->spec.size() != sp.dev_words) { b->smem_reason = "internal: specialised table size"; return; }
    int Lx = b->g.size[0], LxLy = b->g.size[0] * b->g.size[1];
    if (Lx == 1 || LxLy == 1) { b->smem_reason = "degenerate lattice"; return; }
    sp.magic_x = (uint32_t)((0x100000000ull / (uint64_t)Lx) + 1);
    sp.magic_xy = (uint32_t)((0x100000000ull / (uint64_t)LxLy) + 1);
    if (!magic_ok(sp.magic_x, Lx, b->g.ncells) || !magic_ok(sp.magic_xy, LxLy, b->g.ncells)) { b->smem_reason = "no exact reciprocal"; return; }
    sp.n_proc = m->h.n_proc; sp.spuck = m->h.spuck; sp.dim = m->h.dim;
    for (int a = 0; a < 3; ++a) sp.size[a] = b->g.size[a];
    sp.ncells = b->g.ncells; sp.volume = b->g.volume;
    const char* nb = getenv("KMOS_B200_NO_BULK");
    sp.use_bulk = (nb && nb[0] == '1') ? 0 : 1;
    b->smem_ok = true;
}

// Shared-memory kernel vs HBM-resident warp kernel: measured per-warp speeds are about 1 : 0.67 : 0.5
// (all-shared : lists in L2 : everything in HBM with the lattice copy in shared memory; RuO2 20x20 1.21e9 vs
// 9.4e8 at 24 warps per SM each), so the HBM kernel wins when shared memory can host fewer than about half of
// the 24 warps it keeps resident itself (ZGB 64x64: 7 warps per SM, 4.7e8 vs 1.06e9).
static int auto_kernel(const kmos_b200_batch* b) {
    if (b->gen_ok) {
        // the generated kernel keeps its lists in L2 like the interpreter's P1G placement, at about half the
        // instructions per step; the HBM warp kernel wins only when few replicas fit into shared memory
        double hbm_warps = (double)b->R / (double)(b->sm_count > 0 ? b->sm_count : 1);
        if (hbm_warps > 24.0) hbm_warps = 24.0;
        const double gen_score = (double)b->gp.replicas_per_cta * b->gp.ctas_per_sm;
        if (!b->li_ok || 0.4 * hbm_warps <= gen_score) return KMOS_B200_KERNEL_GENERATED;
    }
    if (b->smem_ok && b->li_ok) {
        double hbm_warps = (double)b->R / (double)(b->sm_count > 0 ? b->sm_count : 1);
        if (hbm_warps > 24.0) hbm_warps = 24.0;
        return (0.5 * hbm_warps > b->smem_score) ? KMOS_B200_KERNEL_WARP_HBM : KMOS_B200_KERNEL_SMEM;
    }
    if (b->smem_ok) return KMOS_B200_KERNEL_SMEM;
    return (b->li_ok || b->otf_ok) ? KMOS_B200_KERNEL_WARP_HBM : KMOS_B200_KERNEL_GENERIC;
}

// host -> device copy of a temporary host buffer, ordered on the batch's own (non-blocking) stream
static cudaError_t kb_h2d(kmos_b200_batch* b, void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, b->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(b->stream);
}

static int batch_setup(kmos_b200_batch* b, kmos_b200_model* m, int32_t R, const int32_t size[3]);
static int launch_generic(kmos_b200_batch* b, int mode, long long n, int layer, int only_rep, int32_t* io_proc, int32_t* io_site);

extern "C" int kmos_b200_batch_create(kmos_b200_model* m, int32_t R, const int32_t size[3], int32_t device,
                                      kmos_b200_batch** out) {
    if (!m || !out || R <= 0 || !size) return set_err(KMOS_B200_ERR_ARG, "batch_create: bad arguments");
    int ndev = kmos_b200_device_count();
    if (ndev == 0) return set_err(KMOS_B200_ERR_CUDA, "batch_create: no CUDA device (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err(KMOS_B200_ERR_ARG, "batch_create: bad device index");
    CU(cudaSetDevice(device));
    for (int a = 0; a < m->h.dim; ++a)
        if (size[a] <= 0) return set_err(KMOS_B200_ERR_ARG, "batch_create: bad lattice size");
    kmos_b200_batch* b = new kmos_b200_batch();  // value-initialised: every pointer starts out null
    b->model = m; b->R = R; b->device = device;
    const int rc = batch_setup(b, m, R, size);
    if (rc != KMOS_B200_OK) {  // nothing leaks on an error path: the batch and whatever it allocated so far go
        kmos_b200_batch_destroy(b);
        cudaGetLastError();
        return rc;
    }
    *out = b;
    return KMOS_B200_OK;
}

static int batch_setup(kmos_b200_batch* b, kmos_b200_model* m, int32_t R, const int32_t size[3]) {
    for (int a = 0; a < 3; ++a) b->g.size[a] = a < m->h.dim ? size[a] : 1;
    long long cells = (long long)b->g.size[0] * b->g.size[1] * b->g.size[2];
    if (cells * m->h.spuck > 0x7fffffffLL) return set_err(KMOS_B200_ERR_ARG, "lattice too large");
    b->g.ncells = (int)cells; b->g.volume = (int)cells * m->h.spuck;
    b->idx32 = b->g.ncells >= 65536;
    const int P = m->h.n_proc;
    b->lat_stride = (int)align_up((size_t)b->g.volume, 16);
    b->plane_bytes = align_up((size_t)P * b->g.ncells * (b->idx32 ? 4 : 2), 16);
    CU(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking));
    b->stream = b->own_stream;
    CU(cudaEventCreate(&b->ev0));
    CU(cudaEventCreate(&b->ev1));
    const size_t nblob = m->blob.size() * 4;
    CU(cudaMalloc(&b->d_blob, nblob));
    CU(kb_h2d(b, b->d_blob, m->blob.data(), nblob));
    kb_model_view(m->blob.data(), (int64_t)m->blob.size(), b->d_blob, &b->d);
    const size_t RP = (size_t)R * P;
    CU(cudaMalloc(&b->lattice, (size_t)R * b->lat_stride));
    CU(cudaMalloc(&b->p1, (size_t)R * b->plane_bytes));
    CU(cudaMalloc(&b->p2, (size_t)R * b->plane_bytes));
    CU(cudaMalloc(&b->nsites, RP * 4));
    CU(cudaMalloc(&b->rates, RP * 8));
    CU(cudaMalloc(&b->integ, RP * 8));
    CU(cudaMalloc(&b->accum, RP * 8));
    CU(cudaMalloc(&b->procstat, RP * 8));
    CU(cudaMalloc(&b->sc, (size_t)R * sizeof(KbScalars)));
    CU(cudaMemset(b->lattice, KB_NULL_SPECIES, (size_t)R * b->lat_stride));
    CU(cudaMemset(b->p1, 0, (size_t)R * b->plane_bytes));
    CU(cudaMemset(b->p2, 0, (size_t)R * b->plane_bytes));
    CU(cudaMemset(b->nsites, 0, RP * 4));
    CU(cudaMemset(b->rates, 0, RP * 8));
    CU(cudaMemset(b->integ, 0, RP * 8));
    CU(cudaMemset(b->accum, 0, RP * 8));
    CU(cudaMemset(b->procstat, 0, RP * 8));
    b->rates_matrix = b->accum_proc = b->lut = nullptr;
    if (m->h.backend == KB_BACKEND_OTF) {
        const size_t lut = (size_t)(m->h.lut_total > 0 ? m->h.lut_total : 1);
        CU(cudaMalloc(&b->rates_matrix, RP * (b->g.ncells + 1) * 8));
        CU(cudaMalloc(&b->accum_proc, (size_t)R * b->g.ncells * 8));
        CU(cudaMalloc(&b->lut, (size_t)R * lut * 8));
        CU(cudaMemset(b->rates_matrix, 0, RP * (b->g.ncells + 1) * 8));
        CU(cudaMemset(b->lut, 0, (size_t)R * lut * 8));
    }
    std::vector<KbScalars> sc(R);
    memset(sc.data(), 0, sc.size() * sizeof(KbScalars));
    for (int r = 0; r < R; ++r) { sc[r].seed = 1; sc[r].replica = (uint32_t)r; }
    CU(kb_h2d(b, b->sc, sc.data(), sc.size() * sizeof(KbScalars)));
    b->tally = nullptr; b->occ = nullptr; b->group_of = nullptr; b->tally_groups = 0;
    plan_smem(b);
    plan_latint(b);
    b->image = nullptr;
    b->gen_ok = false; b->d_gen_tab = nullptr; b->d_gen_writes = nullptr; b->d_gen_dev = nullptr; b->gen_image = nullptr;
    b->compact_kind = 0;
    b->compact_valid = false;
    b->d_spec = nullptr;
    b->d_sched = nullptr;
    if (b->smem_ok || b->li_ok) CU(cudaMalloc(&b->d_sched, ((size_t)R + 1) * sizeof(int)));
    if (b->smem_ok) {
        CU(cudaMalloc(&b->image, (size_t)R * b->sp.img_bytes));
        CU(cudaMalloc(&b->d_spec, b->spec.size() * 4));
        CU(kb_h2d(b, b->d_spec, b->spec.data(), b->spec.size() * 4));
    }
    b->kernel = auto_kernel(b);
    // the set-up above used the legacy default stream; the batch's own stream is non-blocking and does not
    // order after it, so everything lands before the first call that uses the stream
    CU(cudaDeviceSynchronize());
    return KMOS_B200_OK;
}

extern "C" void kmos_b200_batch_destroy(kmos_b200_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    cudaFree(b->d_blob); cudaFree(b->lattice); cudaFree(b->p1); cudaFree(b->p2); cudaFree(b->nsites);
    cudaFree(b->rates); cudaFree(b->integ); cudaFree(b->accum); cudaFree(b->procstat); cudaFree(b->sc);
    cudaFree(b->rates_matrix); cudaFree(b->accum_proc); cudaFree(b->lut); cudaFree(b->tally); cudaFree(b->occ); cudaFree(b->group_of); cudaFree(b->image); cudaFree(b->d_spec); cudaFree(b->d_sched);
    cudaFree(b->d_gen_tab); cudaFree(b->d_gen_writes); cudaFree(b->d_gen_dev); cudaFree(b->gen_image);
    if (b->gen.handle) dlclose(b->gen.handle);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    delete b;
}

extern "C" int kmos_b200_batch_volume(const kmos_b200_batch* b) { return b->g.volume; }

extern "C" int kmos_b200_select_kernel(kmos_b200_batch* b, int32_t kind) {
    if (kind == KMOS_B200_KERNEL_AUTO) kind = auto_kernel(b);
    if (kind == KMOS_B200_KERNEL_SMEM && !b->smem_ok)
        return set_err(KMOS_B200_ERR_UNSUPPORTED, "shared-memory kernel unavailable: " + b->smem_reason);
    if (kind == KMOS_B200_KERNEL_WARP_HBM && !b->li_ok && !b->otf_ok)
        return set_err(KMOS_B200_ERR_UNSUPPORTED, "warp-per-replica HBM kernel unavailable for this model/lattice");
    if (kind == KMOS_B200_KERNEL_OTF_FAST) {
        if (!b->otf_ok) return set_err(KMOS_B200_ERR_UNSUPPORTED, "the otf production kernel needs an otf model (up to 64 processes)");
        const long long nblk = ((long long)b->g.ncells + KB_OTFF_BLOCK - 1) >> KB_OTFF_SHIFT;
        if ((long long)b->model->h.n_proc * nblk > (long long)b->g.ncells)
            return set_err(KMOS_B200_ERR_UNSUPPORTED, "otf production kernel: lattice too small for the block sums (P * ceil(ncells/256) > ncells)");
    }
    if (kind == KMOS_B200_KERNEL_GENERATED && !b->gen_ok)
        return set_err(KMOS_B200_ERR_UNSUPPORTED, "no generated proclist attached (kmos_b200_batch_attach_proclist)");
    if (kind != KMOS_B200_KERNEL_SMEM && kind != KMOS_B200_KERNEL_GENERIC && kind != KMOS_B200_KERNEL_WARP_HBM &&
        kind != KMOS_B200_KERNEL_GENERATED && kind != KMOS_B200_KERNEL_OTF_FAST)
        return set_err(KMOS_B200_ERR_ARG, "bad kernel kind");
    b->kernel = kind;
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_kernel_info(kmos_b200_batch* b, int64_t info[12]) {
    memset(info, 0, 12 * sizeof(int64_t));
    info[0] = b->kernel;
    if (b->kernel == KMOS_B200_KERNEL_SMEM) {
        info[1] = b->wpc; info[2] = b->smem_bytes; info[3] = b->ctas_per_sm; info[4] = b->sm_count;
        info[5] = b->sp.rep_bytes; info[6] = (int64_t)b->sp.dev_words * 4;
        info[7] = (b->R + b->wpc - 1) / b->wpc;
        if (info[7] > (int64_t)b->sm_count * b->ctas_per_sm) info[7] = (int64_t)b->sm_count * b->ctas_per_sm;
        info[8] = b->sp.p1_global; info[9] = b->regs; info[10] = b->sp.split; info[11] = b->sp.img_bytes;
    } else if (b->kernel == KMOS_B200_KERNEL_GENERATED) {
        info[1] = b->gp.replicas_per_cta; info[2] = b->gp.smem_bytes; info[3] = b->gp.ctas_per_sm; info[4] = b->gp.sm_count;
        info[5] = b->gp.rep_bytes; info[6] = b->gp.tab_bytes;
        info[7] = (b->R + b->gp.replicas_per_cta - 1) / b->gp.replicas_per_cta;
        if (info[7] > (int64_t)b->gp.sm_count * b->gp.ctas_per_sm) info[7] = (int64_t)b->gp.sm_count * b->gp.ctas_per_sm;
        info[8] = 1; info[9] = b->gp.regs; info[10] = 0; info[11] = b->gp.img_bytes;
    } else if (b->kernel == KMOS_B200_KERNEL_OTF_FAST) {
        info[1] = KB_OTFF_WARPS; info[4] = b->sm_count; info[7] = (b->R + KB_OTFF_WARPS - 1) / KB_OTFF_WARPS; info[8] = 1;
    } else if (b->kernel == KMOS_B200_KERNEL_WARP_HBM && b->otf_ok) {
        info[1] = KB_OTF_WARPS; info[2] = KB_OTF_SMEM; info[4] = b->sm_count;
        info[7] = (b->R + KB_OTF_WARPS - 1) / KB_OTF_WARPS; info[8] = 1;
    } else if (b->kernel == KMOS_B200_KERNEL_WARP_HBM) {
        info[1] = b->li_wpc; info[2] = b->li_smem_bytes; info[4] = b->sm_count; info[5] = b->li.rep_bytes;
        info[6] = (int64_t)b->li.dev_words * 4; info[7] = (b->R + b->li_wpc - 1) / b->li_wpc; info[8] = 1;
    } else {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, b->device) == cudaSuccess) info[4] = prop.multiProcessorCount;
        info[1] = 64; info[7] = (b->R + 63) / 64;
    }
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_synchronize(kmos_b200_batch* b) {
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_batch_set_stream(kmos_b200_batch* b, void* cuda_stream) {
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    b->stream = cuda_stream ? (cudaStream_t)cuda_stream : b->own_stream;
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_timer_start(kmos_b200_batch* b) {
    CU(cudaSetDevice(b->device));
    CU(cudaEventRecord(b->ev0, b->stream));
    return KMOS_B200_OK;
}
extern "C" int kmos_b200_timer_stop(kmos_b200_batch* b, double* ms) {
    CU(cudaSetDevice(b->device));
    CU(cudaEventRecord(b->ev1, b->stream));
    CU(cudaEventSynchronize(b->ev1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, b->ev0, b->ev1));
    if (ms) *ms = f;
    return KMOS_B200_OK;
}

// scalars are edited through a host round trip (setup path, not the step loop)
static int edit_scalars(kmos_b200_batch* b, void (*fn)(KbScalars&, int, const void*), const void* arg) {
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    std::vector<KbScalars> sc(b->R);
    CU(cudaMemcpy(sc.data(), b->sc, sc.size() * sizeof(KbScalars), cudaMemcpyDeviceToHost));
    for (int r = 0; r < b->R; ++r) fn(sc[r], r, arg);
    CU(kb_h2d(b, b->sc, sc.data(), sc.size() * sizeof(KbScalars)));
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_set_seeds(kmos_b200_batch* b, const uint64_t* seeds, const uint32_t* ids) {
    if (!seeds) return set_err(KMOS_B200_ERR_ARG, "set_seeds: null");
    struct A { const uint64_t* s; const uint32_t* i; } a = {seeds, ids};
    return edit_scalars(b, [](KbScalars& s, int r, const void* p) {
        const A* a = (const A*)p;
        s.seed = a->s[r];
        s.replica = a->i ? a->i[r] : (uint32_t)r;
    }, &a);
}

extern "C" int kmos_b200_set_kmc_time(kmos_b200_batch* b, const double* t) {
    if (!t) return set_err(KMOS_B200_ERR_ARG, "set_kmc_time: null");
    return edit_scalars(b, [](KbScalars& s, int r, const void* p) { s.kmc_time = ((const double*)p)[r]; }, t);
}

extern "C" int kmos_b200_set_rates(kmos_b200_batch* b, const double* rates) {
    if (!rates) return set_err(KMOS_B200_ERR_ARG, "set_rates: null");
    const size_t n = (size_t)b->R * b->model->h.n_proc;
    for (size_t i = 0; i < n; ++i)
        if (!(rates[i] >= 0.0)) return set_err(KMOS_B200_ERR_ARG, "set_rates: rate constants must be >= 0");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(b->rates, rates, n * 8, cudaMemcpyHostToDevice, b->stream));
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_set_rate_const(kmos_b200_batch* b, int32_t replica, int32_t proc, double rate) {
    const int P = b->model->h.n_proc;
    if (proc < 1 || proc > P || replica >= b->R || !(rate >= 0.0)) return set_err(KMOS_B200_ERR_ARG, "set_rate_const: bad argument");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    if (replica >= 0) {
        CU(kb_h2d(b, b->rates + (size_t)replica * P + proc - 1, &rate, 8));
    } else {
        std::vector<double> col(b->R, rate);
        CU(cudaMemcpy2DAsync(b->rates + proc - 1, (size_t)P * 8, col.data(), 8, 8, b->R, cudaMemcpyHostToDevice, b->stream));
        CU(cudaStreamSynchronize(b->stream));
    }
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_get_rates(kmos_b200_batch* b, double* out) {
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    CU(cudaMemcpy(out, b->rates, (size_t)b->R * b->model->h.n_proc * 8, cudaMemcpyDeviceToHost));
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_set_otf_lut(kmos_b200_batch* b, const double* lut) {
    if (b->model->h.backend != KB_BACKEND_OTF) return set_err(KMOS_B200_ERR_ARG, "set_otf_lut: not an otf model");
    CU(cudaSetDevice(b->device));
    CU(cudaMemcpyAsync(b->lut, lut, (size_t)b->R * b->model->h.lut_total * 8, cudaMemcpyHostToDevice, b->stream));
    // events already registered keep the rate they were added with: refresh every entry and re-add the rows,
    // as KMC_Model.set_rate_constants does (proclist.recalculate_rates_matrix)
    if (b->initialised) return launch_generic(b, KB_MODE_RECALC, 0, 0, -1, nullptr, nullptr);
    return KMOS_B200_OK;
}

static KbSmemParams smem_params(const kmos_b200_batch* b) {
    KbSmemParams sp = b->sp;
    sp.dev = b->d_spec;
    sp.lattice = b->lattice; sp.nsites = b->nsites; sp.image = b->image;
    sp.p1 = (uint16_t*)b->p1; sp.p2 = (uint16_t*)b->p2;
    sp.rates = b->rates; sp.integ = b->integ; sp.procstat = b->procstat; sp.sc = b->sc;
    sp.R = b->R;
    return sp;
}

// The avail tables live either in the canonical planes p1/p2 (generic engine, getters) or in the compact
// image (shared-memory engine); convert lazily when the other representation is needed.
// the generated kernel's image through the same pack/unpack kernels: every process its own upward list
static KbSmemParams gen_pack_params(const kmos_b200_batch* b) {
    KbSmemParams sp;
    memset(&sp, 0, sizeof sp);
    sp.dev = b->d_gen_dev;
    sp.n_proc = b->model->h.n_proc; sp.ncells = b->g.ncells; sp.cap = b->gp.cap;
    sp.split = 0; sp.off_hi = b->gp.off_p2; sp.off_p2 = b->gp.off_p2; sp.img_bytes = b->gp.img_bytes;
    sp.plane_bytes = (int)b->plane_bytes;
    sp.image = (uint16_t*)b->gen_image; sp.p1 = (uint16_t*)b->p1; sp.p2 = (uint16_t*)b->p2; sp.nsites = b->nsites;
    sp.R = b->R;
    return sp;
}
static int ensure_canonical(kmos_b200_batch* b) {
    if (!b->compact_valid) return KMOS_B200_OK;
    CU(cudaSetDevice(b->device));
    kb_unpack_kernel<<<(b->R + 3) / 4, 128, 0, b->stream>>>(b->compact_kind ? gen_pack_params(b) : smem_params(b));
    CU(cudaGetLastError());
    b->compact_valid = false;
    return KMOS_B200_OK;
}
static int ensure_compact(kmos_b200_batch* b, int kind = 0) {
    if (b->compact_valid && b->compact_kind == kind) return KMOS_B200_OK;
    int rc = ensure_canonical(b);
    if (rc) return rc;
    CU(cudaSetDevice(b->device));
    kb_pack_kernel<<<(b->R + 3) / 4, 128, 0, b->stream>>>(kind ? gen_pack_params(b) : smem_params(b));
    CU(cudaGetLastError());
    b->compact_valid = true;
    b->compact_kind = kind;
    return KMOS_B200_OK;
}

static int launch_generic(kmos_b200_batch* b, int mode, long long n, int layer, int only_rep) {
    return launch_generic(b, mode, n, layer, only_rep, nullptr, nullptr);
}
static int launch_generic(kmos_b200_batch* b, int mode, long long n, int layer, int only_rep,
                          int32_t* io_proc, int32_t* io_site) {
    CU(cudaSetDevice(b->device));
    int rc = ensure_canonical(b);
    if (rc) return rc;
    KbBatchView v = batch_view(b);
    const int threads = 64;
    const int blocks = (b->R + threads - 1) / threads;
    if (b->idx32) kb_generic_kernel<uint32_t><<<blocks, threads, 0, b->stream>>>(v, mode, n, layer, only_rep, io_proc, io_site);
    else kb_generic_kernel<uint16_t><<<blocks, threads, 0, b->stream>>>(v, mode, n, layer, only_rep, io_proc, io_site);
    CU(cudaGetLastError());
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_init_state(kmos_b200_batch* b, int32_t layer) {
    const KbModelView& m = b->model->h;
    if (layer < 0 || layer >= m.n_layers || m.init[2 * layer] < 0) return set_err(KMOS_B200_ERR_ARG, "init_state: bad layer");
    b->initialised = true;
    return launch_generic(b, KB_MODE_INIT, 0, layer, -1);
}

extern "C" int kmos_b200_set_configuration(kmos_b200_batch* b, int32_t replica, const int32_t* species, int32_t layer) {
    const KbModelView& m = b->model->h;
    if (!species || replica >= b->R || layer < 0 || layer >= m.n_layers) return set_err(KMOS_B200_ERR_ARG, "set_configuration: bad argument");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    const int V = b->g.volume;
    const int r0 = replica >= 0 ? replica : 0, r1 = replica >= 0 ? replica + 1 : b->R;
    std::vector<uint8_t> lat((size_t)(r1 - r0) * b->lat_stride, KB_NULL_SPECIES);
    for (int r = r0; r < r1; ++r)
        for (int i = 0; i < V; ++i) {
            int s = species[(size_t)(replica >= 0 ? 0 : r) * V + i];
            if (s >= m.n_species) return set_err(KMOS_B200_ERR_ARG, "set_configuration: species id out of range");
            lat[(size_t)(r - r0) * b->lat_stride + i] = s < 0 ? KB_NULL_SPECIES : (uint8_t)s;
        }
    CU(kb_h2d(b, b->lattice + (size_t)r0 * b->lat_stride, lat.data(), lat.size()));
    b->initialised = true;
    return launch_generic(b, KB_MODE_ADJUST, 0, layer, replica);
}

// proclist.get_next_kmc_step / run_proc_nr for every replica (debugging and replay interface of the reference:
// KMC_Model.get_next_kmc_step / run_proc_nr, kmos/run/__init__.py:1357-1368; tests/test_run/test_run.py)
static int step_io(kmos_b200_batch* b, int mode, int32_t* proc, int32_t* site, bool to_device, bool to_host) {
    if (!proc || !site) return set_err(KMOS_B200_ERR_ARG, "proc/site arrays must not be NULL");
    CU(cudaSetDevice(b->device));
    int32_t* d = nullptr;
    CU(cudaMalloc(&d, (size_t)b->R * 2 * sizeof(int32_t)));
    int rc = KMOS_B200_OK;
    if (to_device) {
        if (cudaMemcpyAsync(d, proc, (size_t)b->R * 4, cudaMemcpyHostToDevice, b->stream) != cudaSuccess ||
            cudaMemcpyAsync(d + b->R, site, (size_t)b->R * 4, cudaMemcpyHostToDevice, b->stream) != cudaSuccess)
            rc = set_err(KMOS_B200_ERR_CUDA, "step_io: host to device copy failed");
    }
    if (!rc) rc = launch_generic(b, mode, 0, 0, -1, d, d + b->R);
    if (!rc && to_host) {
        if (cudaMemcpyAsync(proc, d, (size_t)b->R * 4, cudaMemcpyDeviceToHost, b->stream) != cudaSuccess ||
            cudaMemcpyAsync(site, d + b->R, (size_t)b->R * 4, cudaMemcpyDeviceToHost, b->stream) != cudaSuccess)
            rc = set_err(KMOS_B200_ERR_CUDA, "step_io: device to host copy failed");
    }
    cudaStreamSynchronize(b->stream);
    cudaFree(d);
    return rc;
}

extern "C" int kmos_b200_get_next_kmc_step(kmos_b200_batch* b, int32_t* proc, int32_t* site) {
    return step_io(b, KB_MODE_NEXT, proc, site, false, true);
}

extern "C" int kmos_b200_run_proc_nr(kmos_b200_batch* b, const int32_t* proc, const int32_t* site) {
    return step_io(b, KB_MODE_RUNPROC, const_cast<int32_t*>(proc), const_cast<int32_t*>(site), true, false);
}

// ---------------------------------------------------------------------------------------------------
// generated per-model kernel: attach a proclist module, launch it
// ---------------------------------------------------------------------------------------------------
extern "C" int kmos_b200_batch_attach_proclist(kmos_b200_batch* b, const char* so_path) {
    if (!b || !so_path) return set_err(KMOS_B200_ERR_ARG, "attach_proclist: bad arguments");
    if (b->gen_ok) return set_err(KMOS_B200_ERR_ARG, "attach_proclist: a proclist module is already attached");
    const kmos_b200_model* m = b->model;
    if (m->h.backend != KB_BACKEND_LOCAL_SMART)
        return set_err(KMOS_B200_ERR_UNSUPPORTED, "attach_proclist: generated kernels exist for local_smart models");
    if (b->idx32) return set_err(KMOS_B200_ERR_UNSUPPORTED, "attach_proclist: more than 65535 cells");
    CU(cudaSetDevice(b->device));
    KbGenModule g;
    g.handle = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!g.handle) return set_err(KMOS_B200_ERR_ARG, std::string("attach_proclist: ") + dlerror());
    auto info_fn = (const KbGenInfo* (*)(void))dlsym(g.handle, "kmos_b200_gen_info");
    g.plan = (int (*)(const int32_t*, int32_t, int32_t, KbGenPlan*))dlsym(g.handle, "kmos_b200_gen_plan");
    g.build_tables = (int (*)(const KbGenPlan*, uint32_t*))dlsym(g.handle, "kmos_b200_gen_build_tables");
    g.launch = (int (*)(const KbGenParams*, const KbGenPlan*, int*, int, int, void*))dlsym(g.handle, "kmos_b200_gen_launch");
    auto fail = [&](int code, const std::string& msg) { dlclose(g.handle); return set_err(code, "attach_proclist: " + msg); };
    if (!info_fn || !g.plan || !g.build_tables || !g.launch) return fail(KMOS_B200_ERR_ARG, "not a kmos_b200 proclist module");
    g.info = info_fn();
    if (g.info->abi != KB_GEN_ABI) return fail(KMOS_B200_ERR_MODEL, "module was generated for another ABI version; regenerate it");
    const uint64_t h = kb_gen_fnv1a(m->blob.data(), m->blob.size() * 4);
    if (g.info->model_hash != h || g.info->n_proc != m->h.n_proc || g.info->spuck != m->h.spuck)
        return fail(KMOS_B200_ERR_MODEL, "module was generated from a different model (table hash mismatch)");
    if (g.info->n_classes > 32 || g.info->n_proc > 64) return fail(KMOS_B200_ERR_UNSUPPORTED, "class/process count");
    KbGenPlan gp;
    const int prc = g.plan(b->g.size, b->R, b->device, &gp);
    if (prc != 0) {
        const char* why = prc == -1 ? "lattice has more than 8191 cells" :
                          prc == -2 ? "lattice smaller than twice the interaction range" :
                          prc == -4 ? "one replica does not fit in shared memory" : "device query failed";
        return fail(KMOS_B200_ERR_UNSUPPORTED, why);
    }
    const int P = m->h.n_proc;
    std::vector<uint32_t> tab((size_t)gp.tab_bytes / 4);
    g.build_tables(&gp, tab.data());
    std::vector<int32_t> dev(16 + P, 0);
    dev[9] = 16;
    for (int q = 0; q < P; ++q)
        dev[16 + q] = q | (0 << 6) | ((int)g.info->proc_cls[q] << 7) | ((int)g.info->proc_member[q] << 12);
    uint32_t *d_tab = nullptr, *d_wr = nullptr;
    int32_t* d_dev = nullptr;
    unsigned char* img = nullptr;
    if (cudaMalloc(&d_tab, tab.size() * 4) != cudaSuccess || cudaMalloc(&d_wr, (size_t)P * 16) != cudaSuccess ||
        cudaMalloc(&d_dev, dev.size() * 4) != cudaSuccess || cudaMalloc(&img, (size_t)b->R * gp.img_bytes) != cudaSuccess ||
        cudaMemcpy(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_wr, g.info->writes, (size_t)P * 16, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_dev, dev.data(), dev.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(d_tab); cudaFree(d_wr); cudaFree(d_dev); cudaFree(img);
        return fail(KMOS_B200_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
    }
    if (!b->d_sched && cudaMalloc(&b->d_sched, ((size_t)b->R + 1) * sizeof(int)) != cudaSuccess) {
        cudaFree(d_tab); cudaFree(d_wr); cudaFree(d_dev); cudaFree(img);
        return fail(KMOS_B200_ERR_CUDA, "scheduler counters");
    }
    b->gen = g; b->gp = gp; b->d_gen_tab = d_tab; b->d_gen_writes = d_wr; b->d_gen_dev = d_dev; b->gen_image = img;
    b->gen_ok = true;
    b->sm_count = gp.sm_count;
    b->kernel = auto_kernel(b);
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_batch_detach_proclist(kmos_b200_batch* b) {
    if (!b) return set_err(KMOS_B200_ERR_ARG, "detach_proclist: bad arguments");
    if (!b->gen_ok) return KMOS_B200_OK;
    CU(cudaSetDevice(b->device));
    if (b->compact_valid && b->compact_kind == 1) {
        int rc = ensure_canonical(b);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(b->stream));
    cudaFree(b->d_gen_tab); cudaFree(b->d_gen_writes); cudaFree(b->d_gen_dev); cudaFree(b->gen_image);
    b->d_gen_tab = nullptr; b->d_gen_writes = nullptr; b->d_gen_dev = nullptr; b->gen_image = nullptr;
    if (b->gen.handle) dlclose(b->gen.handle);
    b->gen = KbGenModule();
    b->gen_ok = false;
    b->kernel = auto_kernel(b);
    return KMOS_B200_OK;
}

static int launch_generated(kmos_b200_batch* b, long long n) {
    int rc = ensure_compact(b, 1);
    if (rc) return rc;
    const KbGenPlan& gp = b->gp;
    KbGenParams p;
    memset(&p, 0, sizeof p);
    p.tab = b->d_gen_tab; p.tab_bytes = gp.tab_bytes; p.nbt_off = gp.nbt_off; p.mbar_off = gp.mbar_off;
    p.ncells = gp.ncells; p.cap = gp.cap;
    p.lattice = b->lattice; p.nsites = b->nsites; p.image = b->gen_image; p.rates = b->rates; p.integ = b->integ;
    p.procstat = b->procstat; p.sc = b->sc; p.writes = b->d_gen_writes; p.R = b->R; p.nsteps = n;
    p.rep_bytes = gp.rep_bytes; p.warp_bytes = gp.warp_bytes; p.sm_lat = gp.sm_lat; p.sm_ns = gp.sm_ns;
    p.sm_win = gp.sm_win; p.sm_prod = gp.sm_prod;
    p.stage_off = gp.stage_off; p.stage_bytes = gp.stage_bytes; p.lat_stride = gp.lat_stride;
    p.img_bytes = gp.img_bytes;
    const char* nb = getenv("KMOS_B200_NO_BULK");
    p.use_bulk = (nb && nb[0] == '1') ? 0 : 1;
    // launch geometry (persistent CTAs, epochs) is the module's: it knows how many replicas a warp steps
    const char* wenv = getenv("KMOS_B200_WARPS_PER_CTA");
    const char* ep_env = getenv("KMOS_B200_EPOCHS");
    const int lrc = b->gen.launch(&p, &gp, b->d_sched, wenv ? atoi(wenv) : 0, ep_env ? atoi(ep_env) : 0, (void*)b->stream);
    if (lrc == -1) return set_err(KMOS_B200_ERR_ARG, "do_kmc_steps: too many work items");
    if (lrc != 0) return set_err(KMOS_B200_ERR_CUDA, std::string("generated kernel launch: ") + cudaGetErrorString((cudaError_t)lrc));
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_do_kmc_steps(kmos_b200_batch* b, int64_t n) {
    if (n < 0) return set_err(KMOS_B200_ERR_ARG, "do_kmc_steps: n < 0");
    if (n == 0) return KMOS_B200_OK;
    if (b->kernel == KMOS_B200_KERNEL_GENERIC) return launch_generic(b, KB_MODE_STEPS, n, 0, -1);
    CU(cudaSetDevice(b->device));
    if ((b->kernel == KMOS_B200_KERNEL_WARP_HBM || b->kernel == KMOS_B200_KERNEL_OTF_FAST) && b->otf_ok) {
        int rc0 = ensure_canonical(b);
        if (rc0) return rc0;
        KbOtfParams op;
        op.m = b->d; op.g = b->g; op.R = b->R; op.lat_stride = b->lat_stride;
        op.plane_elems = b->plane_bytes / (b->idx32 ? 4 : 2);
        op.lattice = b->lattice; op.p1 = b->p1; op.p2 = b->p2; op.nsites = b->nsites; op.rates = b->rates;
        op.integ = b->integ; op.accum = b->accum; op.procstat = b->procstat; op.sc = b->sc;
        op.rates_matrix = b->rates_matrix; op.accum_proc = b->accum_proc; op.lut = b->lut; op.nsteps = n;
        if (b->kernel == KMOS_B200_KERNEL_OTF_FAST) {
            const int fblocks = (b->R + KB_OTFF_WARPS - 1) / KB_OTFF_WARPS;
            // lane-parallel events when the model's run_proc routines have the generator's shape (section
            // version 5, devtables.compile_otf_tables); KMOS_B200_OTF_LANES=0 keeps lane 0 interpreting
            const char* le = getenv("KMOS_B200_OTF_LANES");
            const kmos_b200_model* mm = b->model;
            if (!(le && le[0] == '0') && mm->h.dev && mm->h.dev_len >= 16 && mm->h.dev[0] == 6 && mm->h.dev[1] == 1)
                op.lanes = b->d.dev;
            if (b->idx32) kb_otf_fast_kernel<uint32_t><<<fblocks, 32 * KB_OTFF_WARPS, 0, b->stream>>>(op);
            else kb_otf_fast_kernel<uint16_t><<<fblocks, 32 * KB_OTFF_WARPS, 0, b->stream>>>(op);
            CU(cudaGetLastError());
            return KMOS_B200_OK;
        }
        const int threads = 32 * KB_OTF_WARPS, blocks = (b->R + KB_OTF_WARPS - 1) / KB_OTF_WARPS;
        if (b->idx32) {
            CU(cudaFuncSetAttribute(kb_otf_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, KB_OTF_SMEM));
            kb_otf_kernel<uint32_t><<<blocks, threads, KB_OTF_SMEM, b->stream>>>(op);
        } else {
            CU(cudaFuncSetAttribute(kb_otf_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, KB_OTF_SMEM));
            kb_otf_kernel<uint16_t><<<blocks, threads, KB_OTF_SMEM, b->stream>>>(op);
        }
        CU(cudaGetLastError());
        return KMOS_B200_OK;
    }
    if (b->kernel == KMOS_B200_KERNEL_WARP_HBM) {
        int rc0 = ensure_canonical(b);
        if (rc0) return rc0;
        KbLatintParams li = b->li;
        li.dev = b->li_mode ? b->d.dev_hbm : b->d.dev;
        li.lattice = b->lattice; li.lat_stride = b->lat_stride; li.nsites = b->nsites; li.p1 = b->p1; li.p2 = b->p2;
        li.plane_elems = b->plane_bytes / (b->idx32 ? 4 : 2);
        li.rates = b->rates; li.integ = b->integ; li.procstat = b->procstat; li.sc = b->sc; li.R = b->R; li.nsteps = n;
        const int threads = b->li_wpc * 32;
        void (*fn)(const KbLatintParams);
        const int np = b->model->h.n_proc;
        // PPL: processes per lane.  Up to 64 processes: 80 registers, 3 CTAs per SM (the 64-register variant for
        // 4 CTAs spills and measured 2-14 % slower even with more replicas than 24 per SM waiting);
        // 65..256 processes carry 4 or 8 rate/integral/prefix registers per lane and run 2 CTAs per SM.
#define KB_LI(PPLV, IDX, MODEV) \
        (lats ? kb_latint_kernel<PPLV, IDX, MODEV, 3, true> : kb_latint_kernel<PPLV, IDX, MODEV, 3, false>)
#define KB_LI_MODE(IDX, MODEV)                                               \
        (np <= 32 ? KB_LI(1, IDX, MODEV) : np <= 64 ? KB_LI(2, IDX, MODEV)   \
                  : np <= 128 ? kb_latint_kernel<4, IDX, MODEV, 2, false> : kb_latint_kernel<8, IDX, MODEV, 2, false>)
        const bool lats = b->li_lat_bytes > 0;
        if (b->li_mode == 0) fn = b->idx32 ? KB_LI_MODE(uint32_t, 0) : KB_LI_MODE(uint16_t, 0);
        else fn = b->idx32 ? KB_LI_MODE(uint32_t, 1) : KB_LI_MODE(uint16_t, 1);
#undef KB_LI_MODE
#undef KB_LI
        CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, b->li_smem_bytes));
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, b->li_smem_bytes));
        if (per_sm < 1) per_sm = 1;
        int blocks = (b->R + b->li_wpc - 1) / b->li_wpc;
        if (blocks > per_sm * b->sm_count) blocks = per_sm * b->sm_count;
        const long long slots = (long long)blocks * b->li_wpc;
        long long epochs = 1;
        if (b->R > slots) {
            epochs = (24 * slots + b->R - 1) / b->R;
            const long long max_epochs = n / 256 > 0 ? n / 256 : 1;
            if (epochs > max_epochs) epochs = max_epochs;
            if (epochs > 64) epochs = 64;
        }
        const char* ep_env = getenv("KMOS_B200_EPOCHS");
        if (ep_env && atoi(ep_env) > 0) epochs = atoi(ep_env);
        if ((n + epochs - 1) / epochs > 0x40000000LL) epochs = (n + 0x3fffffffLL) / 0x40000000LL;  // 32-bit event counters
        li.chunk = (n + epochs - 1) / epochs;
        epochs = (n + li.chunk - 1) / li.chunk;
        if (epochs * (long long)b->R > 0x7fffffffLL) return set_err(KMOS_B200_ERR_ARG, "do_kmc_steps: too many work items");
        li.n_items = (int)(epochs * b->R);
        li.work_counter = b->d_sched;
        li.done = b->d_sched + 1;
        CU(cudaMemsetAsync(b->d_sched, 0, ((size_t)b->R + 1) * sizeof(int), b->stream));
        fn<<<blocks, threads, b->li_smem_bytes, b->stream>>>(li);
        CU(cudaGetLastError());
        return KMOS_B200_OK;
    }
    if (b->kernel == KMOS_B200_KERNEL_GENERATED) return launch_generated(b, n);
    int rc = ensure_compact(b);
    if (rc) return rc;
    KbSmemParams sp = smem_params(b);
    sp.nsteps = n;
    const int threads = b->wpc * 32;
    // persistent CTAs: as many as stay resident; warps fetch (epoch, replica) items dynamically.  Splitting the
    // n steps into epochs keeps the tail short when the replicas do not divide evenly over the warp slots.
    int blocks = (b->R + b->wpc - 1) / b->wpc;
    const int resident = b->sm_count * b->ctas_per_sm;
    if (blocks > resident) blocks = resident;
    const long long slots = (long long)blocks * b->wpc;
    long long epochs = 1;
    if (b->R > slots) {
        epochs = (24 * slots + b->R - 1) / b->R;          // ~24 items per warp slot
        const long long max_epochs = n / 256 > 0 ? n / 256 : 1;  // at least 256 steps per item
        if (epochs > max_epochs) epochs = max_epochs;
        if (epochs > 64) epochs = 64;
        if (epochs < 1) epochs = 1;
    }
    const char* ep_env = getenv("KMOS_B200_EPOCHS");
    if (ep_env && atoi(ep_env) > 0) epochs = atoi(ep_env);
    // the kernel counts an item's events per process in 32 bits
    if ((n + epochs - 1) / epochs > 0x40000000LL) epochs = (n + 0x3fffffffLL) / 0x40000000LL;
    sp.chunk = (n + epochs - 1) / epochs;
    epochs = (n + sp.chunk - 1) / sp.chunk;
    if (epochs * (long long)b->R > 0x7fffffffLL) return set_err(KMOS_B200_ERR_ARG, "do_kmc_steps: too many work items");
    sp.n_items = (int)(epochs * b->R);
    sp.work_counter = b->d_sched;
    sp.done = b->d_sched + 1;
    CU(cudaMemsetAsync(b->d_sched, 0, ((size_t)b->R + 1) * sizeof(int), b->stream));
    CU(cudaFuncSetAttribute(b->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, b->smem_bytes));
    b->fn<<<blocks, threads, b->smem_bytes, b->stream>>>(sp);
    CU(cudaGetLastError());
    return KMOS_B200_OK;
}

// ---------------------------------------------------------------------------------------------------
// getters
// ---------------------------------------------------------------------------------------------------
template <typename T, typename F>
static int get_scalar(kmos_b200_batch* b, T* out, F f) {
    if (!out) return set_err(KMOS_B200_ERR_ARG, "null output");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    std::vector<KbScalars> sc(b->R);
    CU(cudaMemcpy(sc.data(), b->sc, sc.size() * sizeof(KbScalars), cudaMemcpyDeviceToHost));
    for (int r = 0; r < b->R; ++r) out[r] = f(sc[r]);
    return KMOS_B200_OK;
}
extern "C" int kmos_b200_get_kmc_time(kmos_b200_batch* b, double* out) { return get_scalar(b, out, [](const KbScalars& s) { return s.kmc_time; }); }
extern "C" int kmos_b200_get_kmc_time_step(kmos_b200_batch* b, double* out) { return get_scalar(b, out, [](const KbScalars& s) { return s.kmc_time_step; }); }
extern "C" int kmos_b200_get_kmc_step(kmos_b200_batch* b, int64_t* out) { return get_scalar(b, out, [](const KbScalars& s) { return (int64_t)s.kmc_step; }); }
extern "C" int kmos_b200_get_status(kmos_b200_batch* b, int32_t* out) { return get_scalar(b, out, [](const KbScalars& s) { return s.status; }); }
extern "C" int kmos_b200_get_error_info(kmos_b200_batch* b, int32_t* out) {
    if (!out) return set_err(KMOS_B200_ERR_ARG, "null output");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    std::vector<KbScalars> sc(b->R);
    CU(cudaMemcpy(sc.data(), b->sc, sc.size() * sizeof(KbScalars), cudaMemcpyDeviceToHost));
    for (int r = 0; r < b->R; ++r) memcpy(out + 5 * r, sc[r].err, 20);
    return KMOS_B200_OK;
}

static int get_array(kmos_b200_batch* b, void* out, const void* src, size_t bytes) {
    if (!out) return set_err(KMOS_B200_ERR_ARG, "null output");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    CU(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    return KMOS_B200_OK;
}
extern "C" int kmos_b200_get_procstat(kmos_b200_batch* b, int64_t* out) { return get_array(b, out, b->procstat, (size_t)b->R * b->model->h.n_proc * 8); }
extern "C" int kmos_b200_get_integ_rates(kmos_b200_batch* b, double* out) { return get_array(b, out, b->integ, (size_t)b->R * b->model->h.n_proc * 8); }
extern "C" int kmos_b200_get_nr_of_sites(kmos_b200_batch* b, int32_t* out) { return get_array(b, out, b->nsites, (size_t)b->R * b->model->h.n_proc * 4); }
extern "C" int kmos_b200_get_accum_rates(kmos_b200_batch* b, double* out) {
    int rc = launch_generic(b, KB_MODE_ACCUM, 0, 0, -1);  // base.update_accum_rate, as KMC_Model does before reading
    if (rc) return rc;
    return get_array(b, out, b->accum, (size_t)b->R * b->model->h.n_proc * 8);
}

extern "C" int kmos_b200_get_lattice(kmos_b200_batch* b, int32_t* out) {
    if (!out) return set_err(KMOS_B200_ERR_ARG, "null output");
    std::vector<uint8_t> lat((size_t)b->R * b->lat_stride);
    int rc = get_array(b, lat.data(), b->lattice, lat.size());
    if (rc) return rc;
    const int V = b->g.volume;
    for (int r = 0; r < b->R; ++r)
        for (int i = 0; i < V; ++i) {
            uint8_t s = lat[(size_t)r * b->lat_stride + i];
            out[(size_t)r * V + i] = s == KB_NULL_SPECIES ? -1 : (int32_t)s;
        }
    return KMOS_B200_OK;
}

static int compute_occupation(kmos_b200_batch* b, double* d_out) {
    const KbModelView& m = b->model->h;
    kb_occupation_kernel<<<b->R, 128, m.n_species * m.spuck * sizeof(int), b->stream>>>(
        b->lattice, b->lat_stride, b->R, b->g.volume, m.spuck, m.n_species, b->g.ncells, d_out);
    CU(cudaGetLastError());
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_get_occupation(kmos_b200_batch* b, double* out) {
    if (!out) return set_err(KMOS_B200_ERR_ARG, "null output");
    CU(cudaSetDevice(b->device));
    const KbModelView& m = b->model->h;
    const size_t n = (size_t)b->R * m.n_species * m.spuck;
    double* d = nullptr;
    CU(cudaMalloc(&d, n * 8));
    int rc = compute_occupation(b, d);
    if (!rc) rc = get_array(b, out, d, n * 8);
    cudaFree(d);
    return rc;
}

extern "C" int kmos_b200_get_avail_sites(kmos_b200_batch* b, int32_t replica, int32_t* out) {
    if (!out || replica < 0 || replica >= b->R) return set_err(KMOS_B200_ERR_ARG, "get_avail_sites: bad argument");
    const KbModelView& m = b->model->h;
    const int P = m.n_proc, C = b->g.ncells, V = b->g.volume, sp = m.spuck;
    std::vector<unsigned char> h1(b->plane_bytes), h2(b->plane_bytes);
    std::vector<int32_t> ns(P);
    int rc = ensure_canonical(b);
    if (rc) return rc;
    rc = get_array(b, h1.data(), (const char*)b->p1 + (size_t)replica * b->plane_bytes, b->plane_bytes);
    if (!rc) rc = get_array(b, h2.data(), (const char*)b->p2 + (size_t)replica * b->plane_bytes, b->plane_bytes);
    if (!rc) rc = get_array(b, ns.data(), b->nsites + (size_t)replica * P, (size_t)P * 4);
    if (rc) return rc;
    memset(out, 0, (size_t)P * V * 2 * 4);
    auto at = [&](const std::vector<unsigned char>& h, size_t i) -> int {
        return b->idx32 ? (int)((const uint32_t*)h.data())[i] : (int)((const uint16_t*)h.data())[i];
    };
    for (int q = 0; q < P; ++q) {
        const int n = m.procsite[q];
        for (int k = 0; k < ns[q]; ++k) out[((size_t)q * V + k) * 2] = at(h1, (size_t)q * C + k) * sp + n;
        for (int c = 0; c < C; ++c) out[((size_t)q * V + (size_t)c * sp + n - 1) * 2 + 1] = at(h2, (size_t)q * C + c);
    }
    return KMOS_B200_OK;
}

// base.reload_system for one replica (base.mpy:365-510): every array of the reference's .reload file, written
// verbatim -- in particular avail_sites keeps the stored ORDER (a touch-up would rebuild it in lattice order
// and change the trajectory).  The two planes must describe the same lists; otherwise KMOS_B200_ERR_ARG.
extern "C" int kmos_b200_reload_replica(kmos_b200_batch* b, int32_t replica, const int32_t* species,
                                        const int32_t* avail, const int32_t* nr_of_sites, const int64_t* procstat,
                                        const double* integ_rates, double kmc_time, int64_t kmc_step) {
    if (!species || !avail || !nr_of_sites || !procstat || replica < 0 || replica >= b->R)
        return set_err(KMOS_B200_ERR_ARG, "reload_replica: bad argument");
    const KbModelView& m = b->model->h;
    const int P = m.n_proc, C = b->g.ncells, V = b->g.volume, sp = m.spuck;
    std::vector<uint8_t> lat(b->lat_stride, KB_NULL_SPECIES);
    for (int i = 0; i < V; ++i) {
        if (species[i] >= m.n_species) return set_err(KMOS_B200_ERR_ARG, "reload_replica: species id out of range");
        lat[i] = species[i] < 0 ? KB_NULL_SPECIES : (uint8_t)species[i];
    }
    std::vector<unsigned char> h1(b->plane_bytes, 0), h2(b->plane_bytes, 0);
    auto put = [&](std::vector<unsigned char>& h, size_t i, int v) {
        if (b->idx32) ((uint32_t*)h.data())[i] = (uint32_t)v; else ((uint16_t*)h.data())[i] = (uint16_t)v;
    };
    for (int q = 0; q < P; ++q) {
        const int n = m.procsite[q], nq = nr_of_sites[q];
        if (nq < 0 || nq > C) return set_err(KMOS_B200_ERR_ARG, "reload_replica: nr_of_sites out of range");
        for (int k = 0; k < nq; ++k) {
            const int site = avail[((size_t)q * V + k) * 2];  // avail_sites(q, k, 1), 1-based site number
            if (site < 1 || site > V || (site - 1) % sp + 1 != n || avail[((size_t)q * V + site - 1) * 2 + 1] != k + 1)
                return set_err(KMOS_B200_ERR_ARG, "reload_replica: avail_sites planes are inconsistent");
            put(h1, (size_t)q * C + k, (site - 1) / sp);
            put(h2, (size_t)q * C + (site - 1) / sp, k + 1);
        }
        int back = 0;
        for (int i = 0; i < V; ++i) back += avail[((size_t)q * V + i) * 2 + 1] != 0;
        if (back != nq) return set_err(KMOS_B200_ERR_ARG, "reload_replica: avail_sites planes are inconsistent");
    }
    int rc = ensure_canonical(b);
    if (rc) return rc;
    CU(cudaStreamSynchronize(b->stream));
    std::vector<double> zeros(P, 0.0);
    CU(kb_h2d(b, b->lattice + (size_t)replica * b->lat_stride, lat.data(), lat.size()));
    CU(kb_h2d(b, (char*)b->p1 + (size_t)replica * b->plane_bytes, h1.data(), b->plane_bytes));
    CU(kb_h2d(b, (char*)b->p2 + (size_t)replica * b->plane_bytes, h2.data(), b->plane_bytes));
    CU(kb_h2d(b, b->nsites + (size_t)replica * P, nr_of_sites, (size_t)P * 4));
    CU(kb_h2d(b, b->procstat + (size_t)replica * P, procstat, (size_t)P * 8));
    CU(kb_h2d(b, b->integ + (size_t)replica * P, integ_rates ? integ_rates : zeros.data(), (size_t)P * 8));
    KbScalars sc;
    CU(cudaMemcpy(&sc, b->sc + replica, sizeof sc, cudaMemcpyDeviceToHost));
    sc.kmc_time = kmc_time; sc.kmc_step = kmc_step; sc.kmc_time_step = 0.0; sc.status = KB_OK;
    for (int i = 0; i < 5; ++i) sc.err[i] = 0;
    CU(kb_h2d(b, b->sc + replica, &sc, sizeof sc));
    b->initialised = true;
    if (m.backend == KB_BACKEND_OTF) {
        // The reference's restart file carries no rates_matrix (base_otf.f90:602-663): after reload_system the
        // front-end's set_rate_constants ends with proclist.recalculate_rates_matrix.  Same here: every
        // registered rate is a function of the restored lattice (gr_<proc> table), the rows are re-added.
        CU(cudaMemsetAsync(b->rates_matrix + (size_t)replica * P * (C + 1), 0, (size_t)P * (C + 1) * 8, b->stream));
        return launch_generic(b, KB_MODE_RECALC, 0, 0, replica);
    }
    return launch_generic(b, KB_MODE_ACCUM, 0, 0, replica);
}

extern "C" int kmos_b200_tally_words(const kmos_b200_batch* b) {
    const KbModelView& m = b->model->h;
    return 2 * m.n_proc + m.n_species * m.spuck + 3;
}

extern "C" int kmos_b200_reduce_tallies(kmos_b200_batch* b, const int32_t* group_of, int32_t n_groups, void* dev_out,
                                        double* host_out) {
    if (n_groups <= 0) return set_err(KMOS_B200_ERR_ARG, "reduce_tallies: n_groups <= 0");
    CU(cudaSetDevice(b->device));
    const KbModelView& m = b->model->h;
    const int words = kmos_b200_tally_words(b), nocc = m.n_species * m.spuck;
    if (group_of) {
        for (int r = 0; r < b->R; ++r)
            if (group_of[r] < 0 || group_of[r] >= n_groups) return set_err(KMOS_B200_ERR_ARG, "reduce_tallies: group id out of range");
        if (!b->group_of) CU(cudaMalloc(&b->group_of, (size_t)b->R * 4));
        CU(cudaMemcpyAsync(b->group_of, group_of, (size_t)b->R * 4, cudaMemcpyHostToDevice, b->stream));
    }
    if (!b->occ) CU(cudaMalloc(&b->occ, (size_t)b->R * nocc * 8));
    double* occ = b->occ;
    double* out = (double*)dev_out;
    if (!out) {
        if (b->tally_groups < n_groups) {
            cudaFree(b->tally);
            CU(cudaMalloc(&b->tally, (size_t)n_groups * words * 8));
            b->tally_groups = n_groups;
        }
        out = b->tally;
    }
    int rc = compute_occupation(b, occ);
    if (rc) return rc;
    kb_tally_kernel<<<n_groups, KB_TALLY_THREADS, 0, b->stream>>>(b->sc, b->procstat, b->integ, occ,
                                                            group_of ? b->group_of : nullptr, b->R, m.n_proc, nocc, out, words);
    CU(cudaGetLastError());
    if (host_out) {  // asynchronous for callers that consume the device buffer (NCCL) themselves
        CU(cudaStreamSynchronize(b->stream));
        CU(cudaMemcpy(host_out, out, (size_t)n_groups * words * 8, cudaMemcpyDeviceToHost));
    }
    return KMOS_B200_OK;
}

extern "C" double kmos_b200_philox_next(uint64_t seed, uint32_t replica, uint64_t step, int32_t slot) {
    double t, p, s;
    kb_philox_step(seed, replica, step, &t, &p, &s);
    return slot == 0 ? t : (slot == 1 ? p : s);
}

// ---------------------------------------------------------------------------------------------------
// shared-memory bandwidth microbenchmark (roofline denominator of kb_smem_kernel)
// ---------------------------------------------------------------------------------------------------
__global__ void kb_smem_bw_kernel(uint4* sink, int iters) {
    __shared__ uint4 buf[2048];  // 32 KB
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    int idx = threadIdx.x;
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        const uint4 v = buf[idx & 2047];
        acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
        idx += blockDim.x;
    }
    if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) sink[blockIdx.x] = acc;  // never true; defeats DCE
}

extern "C" int kmos_b200_measure_smem_bandwidth(int32_t device, double* gbps, double* sm_mhz) {
    if (kmos_b200_device_count() == 0) return set_err(KMOS_B200_ERR_CUDA, "no CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 2, threads = 1024, iters = 1 << 15;
    uint4* sink = nullptr;
    CU(cudaMalloc(&sink, (size_t)blocks * sizeof(uint4)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0));
        kb_smem_bw_kernel<<<blocks, threads>>>(sink, iters);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double gb = (double)blocks * threads * iters * 16.0 / 1e9;
        if (rep > 0 && gb / (ms * 1e-3) > best) best = gb / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    if (gbps) *gbps = best;
    if (sm_mhz) *sm_mhz = prop.clockRate / 1000.0;
    return KMOS_B200_OK;
}

#include "kb_fleet.h"  // kmos_b200_fleet_*: the same replicas dealt to several GPUs of this process
