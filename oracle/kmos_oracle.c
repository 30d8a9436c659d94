/*
 * kmos_oracle.c -- CPU restatement of the kmos step loop.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA engine.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it; the product path (kmos_b200/) never does.
 *
 * It follows the reference's Fortran one routine at a time (paths relative to /root/reference):
 *   base:     kmos/fortran_src/base.mpy  (local_smart), base_lat_int.mpy:249,299 (proc==0 no-op),
 *             base_otf.f90 (rates_matrix book-keeping)
 *   lattice:  kmos/fortran_src/lattice.mpy:146-210 (index maps), :343-491 (4-tuple wrappers)
 *   proclist: kmos/fortran_src/proclist_generic_subroutines.mpy:1-110 (step loop), :236-304 (init)
 * The model-specific part of proclist (run_proc_nr, put_/take_/touchup_, run_proc_<p>, nli_<g>, gr_<p>)
 * is not restated by hand: the statements the reference generator wrote are encoded 1:1, in textual
 * order, into byte-code (kmos_b200/tables.py) and interpreted here sequentially, exactly like the
 * compiled Fortran would execute them.
 *
 * Parity pinning: with rng_kind = ORACLE_RNG_GFORTRAN (xoshiro256** + random_seed(put=) scrambling,
 * libgfortran/intrinsics/random.c) this file reproduces the reference's own known-answer trajectory
 * tests/test_run/_tmp_export_{local_smart,lat_int,otf}/ref_procs_sites_*.log for all three backends
 * (tests/test_oracle_golden.py).  The RuO2 / ZGB / pairwise trajectories are not pinned by any reference
 * fixture; for those parity is defined by this restatement under the shared Philox stream.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction: gfortran -O3 on baseline x86-64
 * emits separate mul/add for base.mpy:615-618).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define KB20_MAGIC 0x4B423230
#define KB20_VERSION 4
enum { SEC_ROUTINES = 1, SEC_CODE, SEC_RUNPROC, SEC_INIT, SEC_GR, SEC_PROCSITE, SEC_DEVICE };
enum {
    OP_REPLACE = 1, OP_IF_CAN, OP_DEL, OP_ADD, OP_DEL_NLI, OP_ADD_NLI, OP_ADD_RATE, OP_UPD_RATE, OP_SELECT,
    OP_CASE, OP_DEL_ALL, OP_CALL, OP_RETURN, OP_INC, OP_JUMP
};
enum { BACKEND_LOCAL_SMART = 0, BACKEND_LAT_INT = 1, BACKEND_OTF = 2 };
enum { ORACLE_RNG_PHILOX = 0, ORACLE_RNG_GFORTRAN = 1 };
enum { ORACLE_OK = 0, ORACLE_DEADLOCK = 1, ORACLE_SPECIES_MISMATCH = 2, ORACLE_CAPACITY = 3, ORACLE_BAD_MODEL = 4 };
#define MAX_VARS 8
#define GR_STRIDE (4 + MAX_VARS)

typedef struct {
    /* model */
    int32_t *blob;
    int backend, n_species, n_proc, spuck, dim, default_species, n_layers, default_layer, n_routines, n_gr,
        lut_total;
    int null_species; /* id handed to base.set_null_species by multi-lattice models, else -1 */
    const int32_t *routines, *code, *runproc, *init, *gr;
    /* system (base.mpy:85-205) */
    int size[3];
    int volume;
    int32_t *lattice;     /* lattice(volume)                       */
    int32_t *avail1;      /* avail_sites(proc, k, 1) -> [proc][k]  */
    int32_t *avail2;      /* avail_sites(proc, site, 2)            */
    int32_t *nr_of_sites; /* nr_of_sites(nr_of_proc)               */
    double *rates, *accum_rates, *integ_rates;
    int64_t *procstat;
    double kmc_time, kmc_time_step;
    int64_t kmc_step;
    /* otf (base_otf.f90:152-162) */
    double *rates_matrix;     /* [proc][volume+1]; last column = row total */
    double *accum_rates_proc; /* [volume] */
    double *lut;              /* gr_<proc> values over the finite nr_vars domain, host-supplied */
    /* rng */
    int rng_kind;
    uint32_t philox_key[2];
    uint32_t replica;
    uint64_t xs[4];
    /* status (replaces Fortran `stop`) */
    int status;
    int32_t err[5]; /* old, new, found, site, step -- the post_mortem 5-tuple (run/__init__.py:1492) */
    /* instrumentation for the algorithmic-bytes formula */
    int64_t n_rs, n_chk, n_del, n_gs, n_add, n_upd;
    /* scratch */
    int nr_vars[MAX_VARS];
} oracle_t;

/* ------------------------------------------------------------------------------------------------ */
/* RNG                                                                                               */
/* ------------------------------------------------------------------------------------------------ */

/* Philox4x32-10 (Salmon et al., SC'11).  counter = (step_lo, step_hi, replica, slot), key = seed.    */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The three uniforms of one kMC step.  ran_time in (0,1] (never log(0)); ran_proc, ran_site in [0,1). */
void kmos_oracle_philox_step(uint64_t seed, uint32_t replica, uint64_t step, double out[3]) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c[4] = {(uint32_t)step, (uint32_t)(step >> 32), replica, 0}, a[4], b[4];
    philox4x32_10(c, key, a);
    c[3] = 1;
    philox4x32_10(c, key, b);
    uint64_t x0 = ((uint64_t)a[1] << 32) | a[0], x1 = ((uint64_t)a[3] << 32) | a[2],
             x2 = ((uint64_t)b[1] << 32) | b[0];
    out[0] = (double)((x0 >> 11) + 1) * 0x1.0p-53;
    out[1] = (double)(x1 >> 11) * 0x1.0p-53;
    out[2] = (double)(x2 >> 11) * 0x1.0p-53;
}

/* gfortran >= 7 random_number(): xoshiro256** ; random_seed(put=) xors the user seed with a fixed key
 * (libgfortran/intrinsics/random.c: xor_keys, scramble_seed, prng_next, rnumber_8).                 */
static const uint64_t gf_xor_keys[4] = {0xbd0c5b6e50c2df49ULL, 0xd46061cd46e1df38ULL, 0xbb4f4d4ed6103544ULL,
                                        0x114a583d0756ad39ULL};
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t xoshiro_next(uint64_t s[4]) {
    const uint64_t result = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}
static void gfortran_seed(oracle_t *o, int32_t seed) {
    /* seed_arr(1:seed_size) = seed ; call random_seed(put=seed_arr)   (proclist_generic_subroutines.mpy:253-258);
     * seed_size = 8 int32 = 4 x uint64 for xoshiro256** */
    uint64_t w = ((uint64_t)(uint32_t)seed << 32) | (uint32_t)seed;
    for (int i = 0; i < 4; ++i) o->xs[i] = w ^ gf_xor_keys[i];
}
static double gfortran_random_r8(oracle_t *o) {
    uint64_t v = xoshiro_next(o->xs);
    v &= ~(uint64_t)0 << (64 - 53);
    return (double)v * 0x1.0p-64;
}

static void draw3(oracle_t *o, double r[3]) {
    if (o->rng_kind == ORACLE_RNG_GFORTRAN) {
        r[0] = gfortran_random_r8(o); /* ran_time */
        r[1] = gfortran_random_r8(o); /* ran_proc */
        r[2] = gfortran_random_r8(o); /* ran_site */
    } else {
        uint64_t seed = ((uint64_t)o->philox_key[1] << 32) | o->philox_key[0];
        kmos_oracle_philox_step(seed, o->replica, (uint64_t)o->kmc_step, r);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* lattice (lattice.mpy:146-210)                                                                     */
/* ------------------------------------------------------------------------------------------------ */

static inline int imod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

/* calculate_lattice2nr: 1-based site number of the 4-tuple (x, y, z, n) */
static inline int lattice2nr(const oracle_t *o, const int s[4]) {
    int c = imod(s[0], o->size[0]);
    if (o->dim >= 2) c += o->size[0] * imod(s[1], o->size[1]);
    if (o->dim >= 3) c += o->size[0] * o->size[1] * imod(s[2], o->size[2]);
    return o->spuck * c + s[3];
}
/* calculate_nr2lattice */
static inline void nr2lattice(const oracle_t *o, int nr, int s[4]) {
    int c = (nr - 1) / o->spuck;
    s[3] = nr - o->spuck * c;
    s[0] = c % o->size[0];
    c /= o->size[0];
    s[1] = (o->dim >= 2) ? c % o->size[1] : 0;
    c = (o->dim >= 2) ? c / o->size[1] : 0;
    s[2] = (o->dim >= 3) ? c : 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* base                                                                                              */
/* ------------------------------------------------------------------------------------------------ */

#define AV1(o, p, k) ((o)->avail1[(size_t)((p)-1) * (o)->volume + ((k)-1)])
#define AV2(o, p, s) ((o)->avail2[(size_t)((p)-1) * (o)->volume + ((s)-1)])
#define RM(o, p, k) ((o)->rates_matrix[(size_t)((p)-1) * ((o)->volume + 1) + ((k)-1)])

/* base.mpy:211-265 ; base_lat_int.mpy:249 / base_otf.f90:221-284 (proc==0 no-op, rates row compaction) */
static void del_proc(oracle_t *o, int proc, int site) {
    if (proc <= 0) return;
    o->n_del++;
    int n = o->nr_of_sites[proc - 1];
    int memory_address = AV2(o, proc, site);
    if (memory_address < n) {
        AV1(o, proc, memory_address) = AV1(o, proc, n);
        AV1(o, proc, n) = 0;
        if (o->backend == BACKEND_OTF) {
            RM(o, proc, o->volume + 1) = RM(o, proc, o->volume + 1) - RM(o, proc, memory_address);
            RM(o, proc, memory_address) = RM(o, proc, n);
            RM(o, proc, n) = 0.0;
        }
        AV2(o, proc, AV1(o, proc, memory_address)) = memory_address;
    } else {
        AV1(o, proc, memory_address) = 0;
        if (o->backend == BACKEND_OTF) {
            RM(o, proc, o->volume + 1) = RM(o, proc, o->volume + 1) - RM(o, proc, memory_address);
            RM(o, proc, memory_address) = 0.0;
        }
    }
    AV2(o, proc, site) = 0;
    o->nr_of_sites[proc - 1] = n - 1;
}

/* base.mpy:268-302 ; base_otf.f90:286-331 */
static void add_proc(oracle_t *o, int proc, int site, double rate) {
    if (proc <= 0) return;
    o->n_add++;
    int n = ++o->nr_of_sites[proc - 1];
    AV1(o, proc, n) = site;
    AV2(o, proc, site) = n;
    if (o->backend == BACKEND_OTF) {
        RM(o, proc, o->volume + 1) = RM(o, proc, o->volume + 1) + rate;
        RM(o, proc, n) = rate;
    }
}

/* base_otf.f90:333-364 */
static void update_rates_matrix(oracle_t *o, int proc, int site, double rate) {
    o->n_upd++;
    int memory_address = AV2(o, proc, site);
    RM(o, proc, o->volume + 1) = RM(o, proc, o->volume + 1) + rate - RM(o, proc, memory_address);
    RM(o, proc, memory_address) = rate;
}

/* base.mpy:304-321 */
static inline int can_do(oracle_t *o, int proc, int site) {
    o->n_chk++;
    return AV2(o, proc, site) != 0;
}

/* base.mpy:1187-1231: checked swap; mismatch -> status instead of `stop` */
static void replace_species(oracle_t *o, int site, int old_species, int new_species) {
    o->n_rs++;
    if (old_species != o->lattice[site - 1]) {
        if (o->status == ORACLE_OK) {
            o->status = ORACLE_SPECIES_MISMATCH;
            o->err[0] = old_species; o->err[1] = new_species; o->err[2] = o->lattice[site - 1];
            o->err[3] = site; o->err[4] = (int32_t)o->kmc_step;
        }
        return;
    }
    o->lattice[site - 1] = new_species;
}

/* base.mpy:603-623 ; otf: base_otf.f90:687-717 */
static void update_accum_rate(oracle_t *o) {
    int P = o->n_proc;
    if (o->backend == BACKEND_OTF) {
        for (int i = 1; i <= P; ++i) {
            double tot = 0.0;
            for (int j = 1; j <= o->nr_of_sites[i - 1]; ++j) tot = tot + RM(o, i, j);
            RM(o, i, o->volume + 1) = tot;
            o->accum_rates[i - 1] = (i == 1) ? tot : o->accum_rates[i - 2] + tot;
        }
        return;
    }
    o->accum_rates[0] = o->nr_of_sites[0] * o->rates[0];
    for (int i = 2; i <= P; ++i) o->accum_rates[i - 1] = o->accum_rates[i - 2] + o->nr_of_sites[i - 1] * o->rates[i - 1];
}

/* base.mpy:626-645 ; otf: base_otf.f90:719-739 */
static void update_integ_rate(oracle_t *o) {
    for (int i = 1; i <= o->n_proc; ++i) {
        if (o->backend == BACKEND_OTF)
            o->integ_rates[i - 1] = o->integ_rates[i - 1] + RM(o, i, o->volume + 1) * o->kmc_time_step;
        else
            o->integ_rates[i - 1] = o->integ_rates[i - 1] + o->nr_of_sites[i - 1] * o->rates[i - 1] * o->kmc_time_step;
    }
}

/* base.mpy:1123-1161 (CPU_TIME/walltime dropped) */
static void update_clocks(oracle_t *o, double ran_time) {
    o->kmc_time_step = -log(ran_time) / o->accum_rates[o->n_proc - 1];
    o->kmc_time = o->kmc_time + o->kmc_time_step;
    o->kmc_step = o->kmc_step + 1;
}

/* base.mpy:1234-1338.  arr is 1-based in the comments; returns 1-based index, or 0 on dead-lock.      */
static int interval_search_real(const double *arr, int size, double value) {
    int left = 1, right = size, mid;
    for (;;) {
        mid = (right + left) >> 1;
        if (left >= right) break;
        if (value < arr[mid - 1]) right = mid; else left = mid + 1;
    }
    if (arr[mid - 1] == 0.) {
        for (;;) { /* nonzerosearch: the Fortran walks off the array on an all-zero input */
            if (mid > size) return 0;
            if (arr[mid - 1] > 0.) {
                if (mid >= size) return 0; /* the reference prints its dead-lock message and stops */
                break;
            }
            mid = mid + 1;
        }
    }
    for (;;) { /* leftmostsearch */
        if (mid == 1) break;
        if (arr[mid - 2] >= arr[mid - 1]) mid = mid - 1; else break;
    }
    return mid;
}

/* base.mpy:1075-1120 ; otf: base_otf.f90:1213-1277 */
static int determine_procsite(oracle_t *o, double ran_proc, double ran_site, int *proc, int *site) {
    int P = o->n_proc;
    int p = interval_search_real(o->accum_rates, P, ran_proc * o->accum_rates[P - 1]);
    if (p == 0 || o->nr_of_sites[p - 1] <= 0) { o->status = ORACLE_DEADLOCK; return -1; }
    int n = o->nr_of_sites[p - 1];
    if (o->backend == BACKEND_OTF) {
        o->accum_rates_proc[0] = RM(o, p, 1);
        for (int i = 2; i <= n; ++i) o->accum_rates_proc[i - 1] = o->accum_rates_proc[i - 2] + RM(o, p, i);
        int k = interval_search_real(o->accum_rates_proc, n, ran_site * o->accum_rates_proc[n - 1]);
        if (k == 0) { o->status = ORACLE_DEADLOCK; return -1; }
        *site = AV1(o, p, k);
    } else {
        int k = (int)(1 + ran_site * n); /* int(1+ran_site*(nr_of_sites(proc))) truncates */
        if (k > n) k = n;
        *site = AV1(o, p, k);
    }
    *proc = p;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* proclist: byte-code interpreter for the generated statements                                      */
/* ------------------------------------------------------------------------------------------------ */

static inline int get_species4(oracle_t *o, const int base[4], const int32_t *off) {
    int s[4] = {base[0] + off[0], base[1] + off[1], base[2] + off[2], base[3] + off[3]};
    o->n_gs++;
    return o->lattice[lattice2nr(o, s) - 1];
}
static inline int site_nr(const oracle_t *o, const int base[4], const int32_t *off) {
    int s[4] = {base[0] + off[0], base[1] + off[1], base[2] + off[2], base[3] + off[3]};
    return lattice2nr(o, s);
}

static int exec_routine(oracle_t *o, int rid, const int base[4], int *retval);

static double eval_gr(oracle_t *o, int gid, const int base[4], const int32_t *off) {
    const int32_t *g = o->gr + (size_t)gid * GR_STRIDE;
    int cell[4] = {base[0] + off[0], base[1] + off[1], base[2] + off[2], base[3] + off[3]};
    int saved[MAX_VARS];
    memcpy(saved, o->nr_vars, sizeof saved);
    memset(o->nr_vars, 0, sizeof o->nr_vars);
    int dummy = 0;
    exec_routine(o, g[0], cell, &dummy);
    int idx = 0, stride = 1;
    for (int k = 0; k < g[2]; ++k) { idx += o->nr_vars[k] * stride; stride *= g[4 + k]; }
    memcpy(o->nr_vars, saved, sizeof saved);
    return o->lut[g[3] + idx];
}

/* returns 1 if an OP_RETURN was executed */
static int exec_block(oracle_t *o, const int32_t *pc, const int32_t *end, const int base[4], int *retval) {
    while (pc < end) {
        switch (*pc) {
        case OP_REPLACE:
            replace_species(o, site_nr(o, base, pc + 1), pc[5], pc[6]);
            pc += 7;
            break;
        case OP_IF_CAN: {
            int len = pc[6];
            if (can_do(o, pc[1], site_nr(o, base, pc + 2)))
                if (exec_block(o, pc + 7, pc + 7 + len, base, retval)) return 1;
            pc += 7 + len;
            break;
        }
        case OP_DEL:
            del_proc(o, pc[1], site_nr(o, base, pc + 2));
            pc += 6;
            break;
        case OP_ADD:
            add_proc(o, pc[1], site_nr(o, base, pc + 2), 0.0);
            pc += 6;
            break;
        case OP_DEL_NLI:
        case OP_ADD_NLI: {
            int cell[4] = {base[0] + pc[2], base[1] + pc[3], base[2] + pc[4], base[3] + pc[5]};
            int proc = 0;
            exec_routine(o, pc[1], cell, &proc);
            if (*pc == OP_DEL_NLI) del_proc(o, proc, site_nr(o, base, pc + 6));
            else add_proc(o, proc, site_nr(o, base, pc + 6), 0.0);
            pc += 10;
            break;
        }
        case OP_ADD_RATE:
            add_proc(o, pc[1], site_nr(o, base, pc + 2), eval_gr(o, pc[6], base, pc + 7));
            pc += 11;
            break;
        case OP_UPD_RATE:
            update_rates_matrix(o, pc[1], site_nr(o, base, pc + 2), eval_gr(o, pc[6], base, pc + 7));
            pc += 11;
            break;
        case OP_SELECT: {
            int species = get_species4(o, base, pc + 1);
            int ncases = pc[5], total = pc[6];
            const int32_t *c = pc + 7;
            for (int i = 0; i < ncases; ++i) {
                int32_t mask = c[1];
                int len = c[2];
                int hit = (mask == -1) || (species >= 0 && ((mask >> species) & 1));
                if (hit) {
                    if (exec_block(o, c + 3, c + 3 + len, base, retval)) return 1;
                    break;
                }
                c += 3 + len;
            }
            pc += 7 + total;
            break;
        }
        case OP_DEL_ALL: {
            int site = site_nr(o, base, pc + 1);
            for (int p = 1; p <= o->n_proc; ++p)
                if (can_do(o, p, site)) del_proc(o, p, site);
            pc += 5;
            break;
        }
        case OP_CALL: {
            int nb[4] = {base[0] + pc[2], base[1] + pc[3], base[2] + pc[4], base[3] + pc[5]};
            int dummy = 0;
            exec_routine(o, pc[1], nb, &dummy);
            pc += 6;
            break;
        }
        case OP_RETURN:
            *retval = pc[1];
            return 1;
        case OP_INC:
            o->nr_vars[pc[1]]++;
            pc += 2;
            break;
        case OP_JUMP: /* end of a case body (flat interpreters skip the remaining cases here) */
            pc = end;
            break;
        default:
            o->status = ORACLE_BAD_MODEL;
            return 1;
        }
    }
    return 0;
}

static int exec_routine(oracle_t *o, int rid, const int base[4], int *retval) {
    const int32_t *c = o->code + o->routines[2 * rid];
    return exec_block(o, c, c + o->routines[2 * rid + 1], base, retval);
}

/* run_proc_nr (generated; io/__init__.py:305-465, 1732-1764, 3286-3326) */
static void run_proc_nr(oracle_t *o, int proc, int nr_site) {
    int lsite[4], dummy = 0;
    o->procstat[proc - 1]++; /* increment_procstat, base.mpy:1010-1023 */
    nr2lattice(o, nr_site, lsite);
    exec_routine(o, o->runproc[proc - 1], lsite, &dummy);
}

/* ------------------------------------------------------------------------------------------------ */
/* public API (ctypes)                                                                               */
/* ------------------------------------------------------------------------------------------------ */

static const int32_t *find_section(const int32_t *blob, int id, int *len) {
    int nsec = blob[13];
    for (int i = 0; i < nsec; ++i)
        if (blob[14 + 3 * i] == id) { *len = blob[14 + 3 * i + 2]; return blob + blob[14 + 3 * i + 1]; }
    *len = 0;
    return NULL;
}

void kmos_oracle_destroy(oracle_t *o) {
    if (!o) return;
    free(o->blob); free(o->lattice); free(o->avail1); free(o->avail2); free(o->nr_of_sites); free(o->rates);
    free(o->accum_rates); free(o->integ_rates); free(o->procstat); free(o->rates_matrix);
    free(o->accum_rates_proc); free(o->lut);
    free(o);
}

/* allocate_system (base.mpy:648-744, lattice.mpy:212-317) */
oracle_t *kmos_oracle_create(const int32_t *blob, int64_t n_words, const int32_t size[3]) {
    if (n_words < 14 || blob[0] != KB20_MAGIC || blob[1] != KB20_VERSION) return NULL;
    oracle_t *o = (oracle_t *)calloc(1, sizeof *o);
    o->blob = (int32_t *)malloc((size_t)n_words * 4);
    memcpy(o->blob, blob, (size_t)n_words * 4);
    const int32_t *b = o->blob;
    o->backend = b[2]; o->n_species = b[3]; o->n_proc = b[4]; o->spuck = b[5]; o->dim = b[6];
    o->default_species = b[7] & 0xFFFF; o->null_species = (b[7] >> 16) - 1; o->n_layers = b[8]; o->default_layer = b[9]; o->n_routines = b[10];
    o->n_gr = b[11]; o->lut_total = b[12];
    int len;
    o->routines = find_section(b, SEC_ROUTINES, &len);
    o->code = find_section(b, SEC_CODE, &len);
    o->runproc = find_section(b, SEC_RUNPROC, &len);
    o->init = find_section(b, SEC_INIT, &len);
    o->gr = find_section(b, SEC_GR, &len);
    for (int i = 0; i < 3; ++i) o->size[i] = (i < o->dim) ? size[i] : 1;
    o->volume = o->size[0] * o->size[1] * o->size[2] * o->spuck;
    size_t V = (size_t)o->volume, P = (size_t)o->n_proc;
    o->lattice = (int32_t *)malloc(V * 4);
    for (size_t i = 0; i < V; ++i) o->lattice[i] = o->null_species; /* base.allocate_system: lattice = null_species */
    o->avail1 = (int32_t *)calloc(P * V, 4);
    o->avail2 = (int32_t *)calloc(P * V, 4);
    o->nr_of_sites = (int32_t *)calloc(P, 4);
    o->rates = (double *)calloc(P, 8);
    o->accum_rates = (double *)calloc(P, 8);
    o->integ_rates = (double *)calloc(P, 8);
    o->procstat = (int64_t *)calloc(P, 8);
    if (o->backend == BACKEND_OTF) {
        o->rates_matrix = (double *)calloc(P * (V + 1), 8);
        o->accum_rates_proc = (double *)calloc(V, 8);
        o->lut = (double *)calloc((size_t)(o->lut_total > 0 ? o->lut_total : 1), 8);
    }
    return o;
}

void kmos_oracle_seed(oracle_t *o, int rng_kind, uint64_t seed, uint32_t replica) {
    o->rng_kind = rng_kind;
    o->philox_key[0] = (uint32_t)seed;
    o->philox_key[1] = (uint32_t)(seed >> 32);
    o->replica = replica;
    if (rng_kind == ORACLE_RNG_GFORTRAN) gfortran_seed(o, (int32_t)seed);
}

void kmos_oracle_set_rates(oracle_t *o, const double *rates) { memcpy(o->rates, rates, (size_t)o->n_proc * 8); }
void kmos_oracle_set_lut(oracle_t *o, const double *lut) {
    if (o->lut && o->lut_total > 0) memcpy(o->lut, lut, (size_t)o->lut_total * 8);
}

/* proclist.recalculate_rates_matrix (proclist_generic_subroutines.mpy:307-325): for every process, x outermost,
 * every site that can do it gets the current gr_<proc> value; then base.reaccumulate_rates_matrix
 * (base_otf.f90:366-387) re-adds every row from scratch in memory-address order. */
void kmos_oracle_recalculate_rates_matrix(oracle_t *o) {
    if (o->backend != BACKEND_OTF) return;
    static const int32_t zero_off[4] = {0, 0, 0, 0};
    for (int gid = 0; gid < o->n_gr; ++gid) {
        const int proc = o->gr[(size_t)gid * GR_STRIDE + 1];
        for (int i = 0; i < o->size[0]; ++i)
            for (int j = 0; j < o->size[1]; ++j)
                for (int k = 0; k < o->size[2]; ++k) {
                    int s1[4] = {i, j, k, 1}, base[4] = {i, j, k, 0};
                    const int site = lattice2nr(o, s1);
                    if (can_do(o, proc, site)) update_rates_matrix(o, proc, site, eval_gr(o, gid, base, zero_off));
                }
    }
    for (int proc = 1; proc <= o->n_proc; ++proc) {
        RM(o, proc, o->volume + 1) = 0.0;
        for (int memadd = 1; memadd <= o->volume; ++memadd) {
            if (AV1(o, proc, memadd) > 0) RM(o, proc, o->volume + 1) = RM(o, proc, o->volume + 1) + RM(o, proc, memadd);
            else RM(o, proc, memadd) = 0.0;
        }
    }
}

/* touchup of every cell in the reference's loop order (proclist_generic_subroutines.mpy:279-299) */
static void touchup_all(oracle_t *o, int layer) {
    int rid = o->init[2 * layer + 1], dummy = 0;
    for (int k = 0; k < o->size[2]; ++k)
        for (int j = 0; j < o->size[1]; ++j)
            for (int i = 0; i < o->size[0]; ++i) {
                int base[4] = {i, j, k, 0};
                exec_routine(o, rid, base, &dummy);
            }
}

/* initialize_state (proclist_generic_subroutines.mpy:236-304): reset, default species, touchup */
int kmos_oracle_init_state(oracle_t *o, int layer) {
    if (layer < 0 || layer >= o->n_layers || o->init[2 * layer] < 0) return ORACLE_BAD_MODEL;
    size_t V = (size_t)o->volume, P = (size_t)o->n_proc;
    for (size_t i = 0; i < V; ++i) o->lattice[i] = o->null_species;
    memset(o->avail1, 0, P * V * 4); memset(o->avail2, 0, P * V * 4); memset(o->nr_of_sites, 0, P * 4);
    memset(o->integ_rates, 0, P * 8); memset(o->accum_rates, 0, P * 8); memset(o->procstat, 0, P * 8);
    if (o->rates_matrix) memset(o->rates_matrix, 0, P * (V + 1) * 8);
    o->kmc_time = 0; o->kmc_time_step = 0; o->kmc_step = 0; o->status = ORACLE_OK;
    int dummy = 0;
    for (int k = 0; k < o->size[2]; ++k)
        for (int j = 0; j < o->size[1]; ++j)
            for (int i = 0; i < o->size[0]; ++i) {
                int base[4] = {i, j, k, 0};
                exec_routine(o, o->init[2 * layer], base, &dummy);
            }
    touchup_all(o, layer);
    return o->status;
}

/* KMC_Model._set_configuration + _adjust_database (run/__init__.py:1411-1457): overwrite the lattice
 * (replace_species site by site), then touch up every cell -- x outermost, then y, then z, WITHOUT
 * clearing avail_sites first (touchup strips and re-adds per site, so the previous order matters). */
int kmos_oracle_set_configuration(oracle_t *o, const int32_t *species, int layer) {
    memcpy(o->lattice, species, (size_t)o->volume * 4);
    int rid = o->init[2 * layer + 1], dummy = 0;
    for (int i = 0; i < o->size[0]; ++i)
        for (int j = 0; j < o->size[1]; ++j)
            for (int k = 0; k < o->size[2]; ++k) {
                int base[4] = {i, j, k, 0};
                exec_routine(o, rid, base, &dummy);
            }
    update_accum_rate(o);
    return o->status;
}

/* do_kmc_steps (proclist_generic_subroutines.mpy:1-44) */
int kmos_oracle_do_steps(oracle_t *o, int64_t n) {
    for (int64_t i = 0; i < n && o->status == ORACLE_OK; ++i) {
        double r[3];
        int proc, site;
        draw3(o, r);
        update_accum_rate(o);
        if (!(o->accum_rates[o->n_proc - 1] > 0.)) { o->status = ORACLE_DEADLOCK; break; }
        update_clocks(o, r[0]);
        update_integ_rate(o);
        if (determine_procsite(o, r[1], r[2], &proc, &site)) break;
        run_proc_nr(o, proc, site);
    }
    return o->status;
}

/* get_next_kmc_step (proclist_generic_subroutines.mpy:85-110): note ran_time is passed as the site
 * selector and no clock is advanced -- this is what tests/test_run/test_run.py exercises. */
int kmos_oracle_get_next_kmc_step(oracle_t *o, int32_t *proc, int32_t *site) {
    double r[3];
    draw3(o, r);
    update_accum_rate(o);
    int p = 0, s = 0;
    if (determine_procsite(o, r[1], r[0], &p, &s)) return o->status;
    *proc = p; *site = s;
    return o->status;
}
int kmos_oracle_run_proc_nr(oracle_t *o, int32_t proc, int32_t site) {
    run_proc_nr(o, proc, site);
    return o->status;
}
void kmos_oracle_update_accum_rate(oracle_t *o) { update_accum_rate(o); }
/* base.interval_search_real on a caller-supplied array (unit tests): 1-based index, 0 = the reference's `stop` */
int kmos_oracle_interval_search_real(const double *arr, int32_t size, double value) { return interval_search_real(arr, size, value); }

/* getters */
int kmos_oracle_volume(const oracle_t *o) { return o->volume; }
int kmos_oracle_nproc(const oracle_t *o) { return o->n_proc; }
int kmos_oracle_status(const oracle_t *o, int32_t err[5]) { if (err) memcpy(err, o->err, sizeof o->err); return o->status; }
double kmos_oracle_kmc_time(const oracle_t *o) { return o->kmc_time; }
double kmos_oracle_kmc_time_step(const oracle_t *o) { return o->kmc_time_step; }
void kmos_oracle_set_kmc_time(oracle_t *o, double t) { o->kmc_time = t; } /* base.set_kmc_time */
int64_t kmos_oracle_kmc_step(const oracle_t *o) { return o->kmc_step; }
void kmos_oracle_get_lattice(const oracle_t *o, int32_t *out) { memcpy(out, o->lattice, (size_t)o->volume * 4); }
void kmos_oracle_get_procstat(const oracle_t *o, int64_t *out) { memcpy(out, o->procstat, (size_t)o->n_proc * 8); }
void kmos_oracle_get_nr_of_sites(const oracle_t *o, int32_t *out) { memcpy(out, o->nr_of_sites, (size_t)o->n_proc * 4); }
void kmos_oracle_get_integ_rates(const oracle_t *o, double *out) { memcpy(out, o->integ_rates, (size_t)o->n_proc * 8); }
void kmos_oracle_get_accum_rates(const oracle_t *o, double *out) { memcpy(out, o->accum_rates, (size_t)o->n_proc * 8); }
/* avail_sites as [proc][field][2] int32 (plane 0 = sites, plane 1 = addresses), 1-based contents */
void kmos_oracle_get_avail_sites(const oracle_t *o, int32_t *out) {
    size_t V = (size_t)o->volume;
    for (int p = 0; p < o->n_proc; ++p)
        for (size_t k = 0; k < V; ++k) {
            out[((size_t)p * V + k) * 2 + 0] = o->avail1[(size_t)p * V + k];
            out[((size_t)p * V + k) * 2 + 1] = o->avail2[(size_t)p * V + k];
        }
}
void kmos_oracle_get_rates_matrix_row(const oracle_t *o, int proc, double *out) {
    if (o->rates_matrix) memcpy(out, &RM(o, proc, 1), ((size_t)o->volume + 1) * 8);
}
/* get_occupation (proclist_generic_subroutines.mpy:113-158): out[species][spuck] */
void kmos_oracle_get_occupation(const oracle_t *o, double *out) {
    int ns = o->n_species, sp = o->spuck;
    for (int i = 0; i < ns * sp; ++i) out[i] = 0;
    for (int nr = 1; nr <= o->volume; ++nr) {
        int s = o->lattice[nr - 1];
        if (s >= 0) out[s * sp + ((nr - 1) % sp)] += 1;
    }
    double cells = (double)(o->size[0] * o->size[1] * o->size[2]);
    for (int i = 0; i < ns * sp; ++i) out[i] /= cells;
}
void kmos_oracle_get_counters(const oracle_t *o, int64_t out[6]) {
    out[0] = o->n_rs; out[1] = o->n_chk; out[2] = o->n_del; out[3] = o->n_gs; out[4] = o->n_add; out[5] = o->n_upd;
}
void kmos_oracle_reset_counters(oracle_t *o) { o->n_rs = o->n_chk = o->n_del = o->n_gs = o->n_add = o->n_upd = 0; }
