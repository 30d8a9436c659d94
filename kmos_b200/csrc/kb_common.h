// kb_common.h -- shared definitions of the kmos_b200 engine (host + device).
//
// Reference semantics restated here (paths relative to the kmos checkout):
//   kmos/fortran_src/base.mpy            state arrays (:85-205), add_proc/del_proc (:211-302)
//   kmos/fortran_src/lattice.mpy         index maps (:146-210)
//   kmos/fortran_src/kind_values.f90     iint=int32, ilong=int64, rsingle=rdouble=float64
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KB_HD __host__ __device__ __forceinline__
#define KB_HDN __host__ __device__
#else
#define KB_HD inline
#define KB_HDN
#endif

// Round-to-nearest multiply/add that the compiler may NOT contract into an FMA: the reference's
// accum_rates recurrence is mul-then-add (base.mpy:615-618; gfortran -O3 on baseline x86-64 has no FMA).
#if defined(__CUDA_ARCH__)
#define KB_MUL(a, b) __dmul_rn((a), (b))
#define KB_ADD(a, b) __dadd_rn((a), (b))
#define KB_SUB(a, b) __dsub_rn((a), (b))
#else
// host builds of this header are compiled with -ffp-contract=off
#define KB_MUL(a, b) ((a) * (b))
#define KB_ADD(a, b) ((a) + (b))
#define KB_SUB(a, b) ((a) - (b))
#endif

#define KB20_MAGIC 0x4B423230
#define KB20_VERSION 4

enum { KB_SEC_ROUTINES = 1, KB_SEC_CODE, KB_SEC_RUNPROC, KB_SEC_INIT, KB_SEC_GR, KB_SEC_PROCSITE, KB_SEC_DEVICE,
       KB_SEC_DEVICE_HBM };
enum {
    KB_OP_REPLACE = 1, KB_OP_IF_CAN, KB_OP_DEL, KB_OP_ADD, KB_OP_DEL_NLI, KB_OP_ADD_NLI, KB_OP_ADD_RATE,
    KB_OP_UPD_RATE, KB_OP_SELECT, KB_OP_CASE, KB_OP_DEL_ALL, KB_OP_CALL, KB_OP_RETURN, KB_OP_INC, KB_OP_JUMP
};
enum { KB_BACKEND_LOCAL_SMART = 0, KB_BACKEND_LAT_INT = 1, KB_BACKEND_OTF = 2 };

// per-replica status: replaces the Fortran `stop`s (base.mpy:1205-1228, 1296-1305)
enum {
    KB_OK = 0,
    KB_DEADLOCK = 1,          // total rate 0 / interval_search_real found nothing
    KB_SPECIES_MISMATCH = 2,  // replace_species found another species; err = (old,new,found,site,step)
    KB_CAPACITY = 3,
    KB_BAD_MODEL = 4
};

#define KB_MAX_VARS 8
#define KB_GR_STRIDE (4 + KB_MAX_VARS)
#define KB_NULL_SPECIES 255  // null_species (-1) in the uint8 lattice

// device tables (kmos_b200/devtables.py)
#define KB_DEV_MAX_ROUNDS 8
#define KB_DEV_MAX_WRITES 4
#define KB_DEV_EVENT_STRIDE (4 + KB_DEV_MAX_ROUNDS + 2 * KB_DEV_MAX_WRITES)
#define KB_KIND_DEL_IF 1
#define KB_KIND_ADD 2

struct KbModelView {
    const int32_t* blob;
    int backend, n_species, n_proc, spuck, dim, default_species, n_layers, default_layer, n_routines, n_gr,
        lut_total;
    int null_species;  // id the model passes to base.set_null_species (multi-lattice models), else -1
    const int32_t *routines, *code, *runproc, *init, *gr, *procsite, *dev, *dev_hbm;
    int dev_len, dev_hbm_len;
};

struct KbGeom {
    int size[3];
    int ncells;
    int volume;  // ncells * spuck
};

// One replica's state.  idx_t = uint16_t when ncells < 65536 else uint32_t.
// Every process is registered on exactly one site type (SEC_PROCSITE), so both planes of the
// reference's avail_sites(proc, volume, 2) are stored per *cell*:
//   p1[q][k]  = cell index (0-based) of the k-th available site of process q+1   (k < nsites[q])
//   p2[q][c]  = 1-based position of cell c in p1[q], 0 if process q+1 is not available there
template <typename idx_t>
struct KbReplica {
    uint8_t* lattice;   // [volume]  species id, KB_NULL_SPECIES = null
    int32_t* nsites;    // [P]
    idx_t* p1;          // [P][ncells]
    idx_t* p2;          // [P][ncells]
    const double* rates;  // [P]
    double* integ;      // [P]
    double* accum;      // [P]
    int64_t* procstat;  // [P]
    double* rates_matrix;  // otf: [P][ncells+1]; column ncells = row total (base_otf.f90:152-162)
    double* accum_proc;    // otf: [ncells]
    double* blk = nullptr; // otf production kernel (kb_otf_fast.cuh): [P][blk_n] block sums of rates_matrix rows
    int blk_n = 0;         //   over 256 positions, maintained by add_proc / del_proc / update_rates_matrix
    const double* lut;     // otf: [lut_total]
    double kmc_time, kmc_time_step;
    int64_t kmc_step;
    uint64_t seed;
    uint32_t replica;
    int32_t status;
    int32_t err[5];
};

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (step_lo, step_hi, replica, slot), key = (seed_lo, seed_hi).
// Identical to oracle/kmos_oracle.c so that both sides consume the same stream.
// ---------------------------------------------------------------------------------------------------
KB_HD void kb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                            uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__)
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The three uniforms of kMC step `step`: ran_time in (0,1], ran_proc and ran_site in [0,1).
KB_HD void kb_philox_step(uint64_t seed, uint32_t replica, uint64_t step, double* ran_time, double* ran_proc,
                          double* ran_site) {
    uint32_t a[4], b[4];
    kb_philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), replica, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), a);
    kb_philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), replica, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), b);
    uint64_t x0 = ((uint64_t)a[1] << 32) | a[0], x1 = ((uint64_t)a[3] << 32) | a[2], x2 = ((uint64_t)b[1] << 32) | b[0];
    *ran_time = (double)((x0 >> 11) + 1) * 0x1.0p-53;
    *ran_proc = (double)(x1 >> 11) * 0x1.0p-53;
    *ran_site = (double)(x2 >> 11) * 0x1.0p-53;
}

KB_HD int kb_imod(int a, int n) {
    int r = a % n;
    return r < 0 ? r + n : r;
}

// lattice.mpy:146-168 calculate_lattice2nr, split into cell index (0-based) and site type (1-based)
// a mod n for a coordinate plus a small offset: one conditional add or subtract in the usual case (the reference's
// lattice2nr table covers [-L, 2L) and nothing else, lattice.mpy:121-132), the division only beyond it
KB_HD int kb_wrap(int a, int n) {
    if ((unsigned)a >= (unsigned)n) {
        a = a < 0 ? a + n : a - n;
        if ((unsigned)a >= (unsigned)n) a = kb_imod(a, n);
    }
    return a;
}

KB_HD int kb_cell_of(const KbModelView& m, const KbGeom& g, int x, int y, int z) {
    int c = kb_wrap(x, g.size[0]);
    if (m.dim >= 2) c += g.size[0] * kb_wrap(y, g.size[1]);
    if (m.dim >= 3) c += g.size[0] * g.size[1] * kb_wrap(z, g.size[2]);
    return c;
}

static inline const int32_t* kb_find_section(const int32_t* blob, int id, int* len) {
    int nsec = blob[13];
    for (int i = 0; i < nsec; ++i)
        if (blob[14 + 3 * i] == id) {
            *len = blob[14 + 3 * i + 2];
            return blob + blob[14 + 3 * i + 1];
        }
    *len = 0;
    return nullptr;
}

// Fill a model view whose pointers are relative to `base` (host copy or device copy of the blob),
// reading the header from the host copy `hblob`.
static inline bool kb_model_view(const int32_t* hblob, int64_t n_words, const int32_t* base, KbModelView* m) {
    if (n_words < 14 || hblob[0] != KB20_MAGIC || hblob[1] != KB20_VERSION) return false;
    // a malformed blob must not send the section pointers out of bounds: table and every section inside n_words
    const int64_t nsec = hblob[13];
    if (nsec < 0 || 14 + 3 * nsec > n_words) return false;
    for (int64_t i = 0; i < nsec; ++i) {
        const int64_t off = hblob[14 + 3 * i + 1], len = hblob[14 + 3 * i + 2];
        if (off < 0 || len < 0 || off + len > n_words) return false;
    }
    m->blob = base;
    m->backend = hblob[2]; m->n_species = hblob[3]; m->n_proc = hblob[4]; m->spuck = hblob[5]; m->dim = hblob[6];
    m->default_species = hblob[7] & 0xFFFF; m->null_species = (hblob[7] >> 16) - 1; m->n_layers = hblob[8]; m->default_layer = hblob[9]; m->n_routines = hblob[10];
    m->n_gr = hblob[11]; m->lut_total = hblob[12];
    int len;
    const int32_t* p;
#define KB_SEC(field, id)                    \
    p = kb_find_section(hblob, id, &len);    \
    m->field = p ? base + (p - hblob) : nullptr;
    KB_SEC(routines, KB_SEC_ROUTINES)
    KB_SEC(code, KB_SEC_CODE)
    KB_SEC(runproc, KB_SEC_RUNPROC)
    KB_SEC(init, KB_SEC_INIT)
    KB_SEC(gr, KB_SEC_GR)
    KB_SEC(procsite, KB_SEC_PROCSITE)
    KB_SEC(dev, KB_SEC_DEVICE)
    m->dev_len = len;
    KB_SEC(dev_hbm, KB_SEC_DEVICE_HBM)
    m->dev_hbm_len = len;
#undef KB_SEC
    return m->routines && m->code && m->runproc && m->init && m->procsite;
}
