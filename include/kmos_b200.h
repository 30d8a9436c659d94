/*
 * kmos_b200.h -- C-ABI of the B200-native batched kMC step engine (libkmos_b200.so).
 *
 * Drop-in boundary for ONE path of kmos: the per-step event cycle of the generated base/lattice/proclist
 * Fortran modules.  Every entry point names the reference interface it replaces (paths relative to the
 * kmos checkout); INTEGRATION.md shows the f2py-shaped Python shim and the ISO_C_BINDING interface block
 * that bind them.  Conventions kept from the reference at this boundary:
 *   - process numbers are 1-based, in process_list order (kmos/fortran_src/proclist_constants.mpy:67-81)
 *   - species ids are 0-based, sorted by name; null_species = -1
 *   - site numbers are 1-based: nr = spuck*(x + Lx*(y + Ly*z)) + n   (kmos/fortran_src/lattice.mpy:146-168)
 *   - kinds: iint=int32_t, ilong=int64_t, rsingle=rdouble=double (kmos/fortran_src/kind_values.f90:10-15)
 * Differences: state is owned by opaque handles (the reference keeps Fortran module globals, one model per
 * process: kmos/run/__init__.py:150-153); a batch holds R independent replicas; errors are return codes and
 * per-replica status words instead of `stop`.
 *
 * All functions return KMOS_B200_OK (0) or a negative error code; kmos_b200_last_error() gives the text.
 * Plain pointers and sizes only; caller owns every output buffer.  There is no CPU fallback: every call
 * that touches state needs a CUDA device.
 */
#ifndef KMOS_B200_H
#define KMOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kmos_b200_model kmos_b200_model;
typedef struct kmos_b200_batch kmos_b200_batch;

enum {
    KMOS_B200_OK = 0,
    KMOS_B200_ERR_ARG = -1,
    KMOS_B200_ERR_MODEL = -2,       /* malformed / unsupported model tables */
    KMOS_B200_ERR_CUDA = -3,
    KMOS_B200_ERR_UNSUPPORTED = -4
};

/* per-replica status (kmos_b200_get_status); replaces the reference's print + stop */
enum {
    KMOS_B200_REPLICA_OK = 0,
    KMOS_B200_REPLICA_DEADLOCK = 1,          /* base.mpy:1296-1305: no process available */
    KMOS_B200_REPLICA_SPECIES_MISMATCH = 2,  /* base.mpy:1205-1228: replace_species found another species */
    KMOS_B200_REPLICA_CAPACITY = 3,
    KMOS_B200_REPLICA_BAD_MODEL = 4
};

/* which stepping kernel a batch uses */
enum {
    KMOS_B200_KERNEL_AUTO = 0,
    KMOS_B200_KERNEL_GENERIC = 1, /* thread-per-replica byte-code engine, state in HBM (all backends) */
    KMOS_B200_KERNEL_SMEM = 2,    /* warp-per-replica, state in shared memory (local_smart that fits) */
    KMOS_B200_KERNEL_WARP_HBM = 3, /* warp-per-replica, state in HBM/L2 (lat_int: nli_* decision trees per lane) */
    KMOS_B200_KERNEL_GENERATED = 4, /* exporter-generated per-model CUDA (proclist_<model>.cu), local_smart */
    KMOS_B200_KERNEL_OTF_FAST = 5  /* otf production mode: sub-linear event selection over block sums of
                                    * rates_matrix (the O(log N) replacement the reference announces,
                                    * doc/source/topic_guides/otf_backend.rst:198-202).  Same distribution, same
                                    * prefix order, NOT bit-exact (floating-point association); never chosen by
                                    * KMOS_B200_KERNEL_AUTO */
};

const char *kmos_b200_last_error(void);

/* Number of CUDA devices visible (0 if none / no driver). */
int kmos_b200_device_count(void);

/* ---- model: the rule tables the kmos exporter emits next to the Fortran ----------------------------
 * replaces: the compiled proclist module (kmos/io/__init__.py:3884-3974 export_source -> f2py build,
 * kmos/utils/__init__.py:406-521).  `blob` is the int32 table image written by kmos_b200.tables. */
int kmos_b200_model_create(const int32_t *blob, int64_t n_words, kmos_b200_model **out);
void kmos_b200_model_destroy(kmos_b200_model *m);
int kmos_b200_model_nproc(const kmos_b200_model *m);     /* proclist.nr_of_proc */
int kmos_b200_model_nspecies(const kmos_b200_model *m);  /* proclist.nr_of_species */
int kmos_b200_model_spuck(const kmos_b200_model *m);     /* lattice.spuck */
int kmos_b200_model_lut_size(const kmos_b200_model *m);  /* otf: doubles per replica in set_otf_lut */

/* ---- batch: R replicas of one model on one GPU ------------------------------------------------------
 * replaces: lattice.allocate_system / base.allocate_system (lattice.mpy:212-317, base.mpy:648-744) as
 * called by proclist.init(size, name, layer, seed, no_banner) (proclist_generic_subroutines.mpy:166-230).
 * size[3]: unit cells per axis (entries beyond model_dimension ignored). */
int kmos_b200_batch_create(kmos_b200_model *m, int32_t n_replicas, const int32_t size[3], int32_t device,
                           kmos_b200_batch **out);
void kmos_b200_batch_destroy(kmos_b200_batch *b);
int kmos_b200_batch_volume(const kmos_b200_batch *b); /* base.get_volume */
int kmos_b200_select_kernel(kmos_b200_batch *b, int32_t kind);
/* info[0]=kernel in use, [1]=replicas per CTA, [2]=dynamic smem bytes per CTA, [3]=CTAs per SM,
 * [4]=SM count, [5]=bytes of state per replica in shared memory, [6]=device table bytes, [7]=grid size,
 * [8]=1 if the avail-site lists stay in HBM/L2, [9]=registers per thread, [10]=1 if split list storage,
 * [11]=bytes of the compact avail image per replica */
int kmos_b200_kernel_info(kmos_b200_batch *b, int64_t info[12]);

/* Attach the model's generated proclist (proclist_<model>_<hash>.so, written by kmos_b200.codegen next to the
 * Fortran and compiled with nvcc for sm_100a) to the batch.  replaces: linking the compiled, model-specific
 * proclist.f90 -- run_proc_nr and its put_/take_ routines (kmos/io/__init__.py:305-465, 2219-2409) -- into
 * kmc_model (kmos/utils/__init__.py:406-521).  The module must have been generated from the same model blob
 * (hash checked).  Afterwards KMOS_B200_KERNEL_AUTO prefers KMOS_B200_KERNEL_GENERATED where it fits; the
 * table interpreter kernels stay available through kmos_b200_select_kernel. */
int kmos_b200_batch_attach_proclist(kmos_b200_batch *b, const char *so_path);
/* Unload the attached module again (state is converted back to the canonical planes first); the batch returns
 * to the table-driven kernels.  Lets a front-end try the module of another lane-group width: how many replicas
 * a width keeps resident depends on the lattice, which is only known when the batch exists. */
int kmos_b200_batch_detach_proclist(kmos_b200_batch *b);

/* RNG: per-replica Philox4x32-10 stream, key = seed, counter = (kmc_step, replica_id, slot).
 * replaces: random_seed(put=seed_arr) in initialize_state (proclist_generic_subroutines.mpy:253-258).
 * seeds[R]; replica_ids[R] may be NULL (then 0..R-1). */
int kmos_b200_set_seeds(kmos_b200_batch *b, const uint64_t *seeds, const uint32_t *replica_ids);

/* base.set_rate_const(proc, rate) (base.mpy:581-600) for the whole batch: rates[R][P] (host pointer). */
int kmos_b200_set_rates(kmos_b200_batch *b, const double *rates);
/* same, one entry; replica = -1 broadcasts */
int kmos_b200_set_rate_const(kmos_b200_batch *b, int32_t replica, int32_t proc, double rate);
/* base.get_rate(proc) */
int kmos_b200_get_rates(kmos_b200_batch *b, double *rates);
/* otf: tabulated gr_<proc> values, lut[R][lut_size]; replaces proclist_pars.update_user_parameter /
 * update_chempot + recalculate_rates_matrix (proclist_generic_subroutines.mpy:307-325) */
int kmos_b200_set_otf_lut(kmos_b200_batch *b, const double *lut);

/* proclist.initialize_state(layer, seed) minus the seeding (proclist_generic_subroutines.mpy:236-304):
 * null everything, default species, touchup every cell in the reference's loop order. */
int kmos_b200_init_state(kmos_b200_batch *b, int32_t layer);
/* KMC_Model._set_configuration + _adjust_database (kmos/run/__init__.py:1411-1457).
 * species[V] (replica >= 0) or species[R][V] (replica = -1), site-number order. */
int kmos_b200_set_configuration(kmos_b200_batch *b, int32_t replica, const int32_t *species, int32_t layer);

/* base.reload_system (base.mpy:365-510) for one replica: the arrays of the reference's <system_name>.reload file
 * (written by base.save_system, :517-578) restored verbatim -- lattice[V], avail_sites[P][V][2] in the layout
 * of kmos_b200_get_avail_sites (order preserved, no touch-up), nr_of_sites[P], procstat[P], kmc_time, kmc_step
 * -- plus integ_rates[P] (NULL = zeros; the reference's file forgets them).  The Philox stream continues at
 * kmc_step.  otf: the file carries no rates_matrix (nor does the reference's, base_otf.f90:602-663); the rows are
 * rebuilt from the restored lattice and the current gr_<proc> table (proclist.recalculate_rates_matrix). */
int kmos_b200_reload_replica(kmos_b200_batch *b, int32_t replica, const int32_t *species, const int32_t *avail_sites,
                             const int32_t *nr_of_sites, const int64_t *procstat, const double *integ_rates,
                             double kmc_time, int64_t kmc_step);

/* proclist.do_kmc_steps(n) (proclist_generic_subroutines.mpy:1-44) on every replica.  Asynchronous on the
 * batch's stream; getters synchronise. */
int kmos_b200_do_kmc_steps(kmos_b200_batch *b, int64_t n);
int kmos_b200_synchronize(kmos_b200_batch *b);
/* proclist.get_next_kmc_step(proc, site) (proclist_generic_subroutines.mpy:85-110) for every replica: the next
 * step's (process, site number), both 1-based, without executing it or advancing the clock; as in the
 * reference the site is selected with ran_time.  proc[R], site[R]; 0/0 for a dead-locked replica. */
int kmos_b200_get_next_kmc_step(kmos_b200_batch *b, int32_t *proc, int32_t *site);
/* proclist.run_proc_nr(proc, nr_site) (generated; kmos/io/__init__.py:305-465): increment_procstat + the
 * event's rule code on replica r for (proc[r], site[r]); proc[r] = 0 skips the replica.  No random numbers, no
 * clock update (KMC_Model.run_proc_nr, kmos/run/__init__.py:1357; replay loop of tests/test_run/test_run.py). */
int kmos_b200_run_proc_nr(kmos_b200_batch *b, const int32_t *proc, const int32_t *site);
/* Run the batch on a caller-owned CUDA stream (cudaStream_t as void*, e.g. torch's current stream) so that
 * the caller's events and collectives order against the engine's kernels.  NULL restores the own stream. */
int kmos_b200_batch_set_stream(kmos_b200_batch *b, void *cuda_stream);
/* CUDA-event timing on the batch's stream: start/stop bracket any sequence of calls */
int kmos_b200_timer_start(kmos_b200_batch *b);
int kmos_b200_timer_stop(kmos_b200_batch *b, double *milliseconds);

/* ---- observables (caller-owned host buffers) ---------------------------------------------------------- */
int kmos_b200_get_kmc_time(kmos_b200_batch *b, double *out /*[R]*/);       /* base.get_kmc_time      */
int kmos_b200_get_kmc_time_step(kmos_b200_batch *b, double *out /*[R]*/);  /* base.get_kmc_time_step */
int kmos_b200_get_kmc_step(kmos_b200_batch *b, int64_t *out /*[R]*/);      /* base.get_kmc_step      */
int kmos_b200_set_kmc_time(kmos_b200_batch *b, const double *t /*[R]*/);   /* base.set_kmc_time      */
int kmos_b200_get_procstat(kmos_b200_batch *b, int64_t *out /*[R][P]*/);   /* base.get_procstat      */
int kmos_b200_get_integ_rates(kmos_b200_batch *b, double *out /*[R][P]*/); /* base.get_integ_rate    */
int kmos_b200_get_nr_of_sites(kmos_b200_batch *b, int32_t *out /*[R][P]*/);/* base.get_nrofsites     */
int kmos_b200_get_accum_rates(kmos_b200_batch *b, double *out /*[R][P]*/); /* base.get_accum_rate (after update_accum_rate) */
int kmos_b200_get_lattice(kmos_b200_batch *b, int32_t *out /*[R][V]*/);    /* lattice.get_species per site */
/* proclist.get_occupation (proclist_generic_subroutines.mpy:113-158): out[R][n_species][spuck] */
int kmos_b200_get_occupation(kmos_b200_batch *b, double *out);
/* base.get_avail_site for one replica, whole array: out[P][V][2] int32, 1-based like avail_sites(:,:,:) */
int kmos_b200_get_avail_sites(kmos_b200_batch *b, int32_t replica, int32_t *out);
int kmos_b200_get_status(kmos_b200_batch *b, int32_t *out /*[R]*/);
/* (old, new, found, site, step) of the first species mismatch: KMC_Model.post_mortem's err_code */
int kmos_b200_get_error_info(kmos_b200_batch *b, int32_t *out /*[R][5]*/);

/* ---- tallies for the multi-GPU reduce ------------------------------------------------------
 * Sum over the batch's replicas, grouped: group_of[R] in [0, n_groups).  Layout per group (doubles):
 *   [P] procstat (exact below 2^53) | [P] integ_rates | [n_species*spuck] occupation | kmc_time | kmc_steps | n_replicas
 * `dev_out` is a DEVICE pointer (e.g. a torch tensor's data_ptr) of n_groups*tally_words doubles so that the
 * caller can hand it straight to ncclAllReduce; host_out (may be NULL) receives a copy. */
int kmos_b200_tally_words(const kmos_b200_batch *b);
int kmos_b200_reduce_tallies(kmos_b200_batch *b, const int32_t *group_of, int32_t n_groups, void *dev_out,
                             double *host_out);

/* ---- fleet: R replicas of one model sharded over several GPUs of ONE process -----------------------------
 * replaces: nothing in the reference (its one trajectory per process is spread by ModelRunner.run(cores=N),
 * kmos/run/__init__.py:2330-2366, a multiprocessing pool); SURVEY 8b's `gpu_ids[], n_gpus` arguments.  For callers
 * that are not launched one process per GPU -- the Fortran templates over ISO_C_BINDING, a plain Python session.
 * Shard k of n_gpus holds replicas [R*k/n_gpus, R*(k+1)/n_gpus) as an ordinary batch on gpu_ids[k]; a device may
 * be named more than once.  seeds[R] (NULL: 0..R-1) are Philox keys, the counter carries the global replica
 * number, so trajectories do not depend on n_gpus.  do_kmc_steps enqueues on every GPU and returns; getters take
 * [R]-leading host buffers exactly like the batch getters and synchronise.  Everything else (set_configuration,
 * avail_sites, restart, ...) goes through kmos_b200_fleet_shard(k) and the batch API. */
typedef struct kmos_b200_fleet kmos_b200_fleet;
int kmos_b200_fleet_create(kmos_b200_model *m, int32_t n_replicas, const int32_t size[3], const uint64_t *seeds,
                           const int32_t *gpu_ids, int32_t n_gpus, kmos_b200_fleet **out);
void kmos_b200_fleet_destroy(kmos_b200_fleet *f);
int kmos_b200_fleet_n_shards(const kmos_b200_fleet *f); /* <= n_gpus: GPUs without a replica hold no shard */
kmos_b200_batch *kmos_b200_fleet_shard(kmos_b200_fleet *f, int32_t k, int32_t *first_replica, int32_t *n_replicas);
int kmos_b200_fleet_attach_proclist(kmos_b200_fleet *f, const char *so_path);
int kmos_b200_fleet_select_kernel(kmos_b200_fleet *f, int32_t kind);
int kmos_b200_fleet_set_rates(kmos_b200_fleet *f, const double *rates /*[R][P]*/);
int kmos_b200_fleet_set_otf_lut(kmos_b200_fleet *f, const double *lut /*[R][lut_size]*/);
int kmos_b200_fleet_init_state(kmos_b200_fleet *f, int32_t layer);
int kmos_b200_fleet_do_kmc_steps(kmos_b200_fleet *f, int64_t n);
int kmos_b200_fleet_synchronize(kmos_b200_fleet *f);
int kmos_b200_fleet_get_kmc_time(kmos_b200_fleet *f, double *out /*[R]*/);
int kmos_b200_fleet_get_kmc_step(kmos_b200_fleet *f, int64_t *out /*[R]*/);
int kmos_b200_fleet_get_status(kmos_b200_fleet *f, int32_t *out /*[R]*/);
int kmos_b200_fleet_get_procstat(kmos_b200_fleet *f, int64_t *out /*[R][P]*/);
int kmos_b200_fleet_get_integ_rates(kmos_b200_fleet *f, double *out /*[R][P]*/);
int kmos_b200_fleet_get_nr_of_sites(kmos_b200_fleet *f, int32_t *out /*[R][P]*/);
int kmos_b200_fleet_get_lattice(kmos_b200_fleet *f, int32_t *out /*[R][V]*/);
int kmos_b200_fleet_get_occupation(kmos_b200_fleet *f, double *out /*[R][n_species][spuck]*/);
/* kmos_b200_reduce_tallies over all shards: group_of[R] (NULL: one group), host_out[n_groups][tally_words];
 * per-GPU device reductions, partial sums added on the host in shard order. */
int kmos_b200_fleet_reduce_tallies(kmos_b200_fleet *f, const int32_t *group_of, int32_t n_groups, double *host_out);

/* Validation hook shared with a Fortran validation build: the uniform number `slot` (0 ran_time,
 * 1 ran_proc, 2 ran_site) of kMC step `step` -- what `call random_number(x)` is rewritten to
 * (pattern: kmos/utils/__init__.py:818-829). */
double kmos_b200_philox_next(uint64_t seed, uint32_t replica, uint64_t step, int32_t slot);

/* Measured shared-memory read bandwidth of `device` in GB/s (LDS.128 streaming microbenchmark, all SMs):
 * the roofline denominator for the shared-memory step kernel. */
int kmos_b200_measure_smem_bandwidth(int32_t device, double *gb_per_s, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif
