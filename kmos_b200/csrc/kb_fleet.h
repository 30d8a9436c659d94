// kb_fleet.h -- R replicas of one model dealt to several GPUs of ONE process (`gpu_ids[], n_gpus`).
//
// The batch API above is one GPU per handle; under torchrun every rank owns one batch.  A caller that cannot be
// launched that way -- the Fortran templates through ISO_C_BINDING, a plain `python` session driving KMC_Model --
// gets the same sharding here: shard k of G holds the replicas [R*k/G, R*(k+1)/G) (kmos_b200/parallel.py
// shard_bounds), each shard is an ordinary kmos_b200_batch on gpu_ids[k] with its own stream, do_kmc_steps enqueues
// on every shard before anything waits, getters concatenate in replica order.  There is no inter-GPU traffic while
// stepping (replicas are independent trajectories); the tally reduce sums the shards' partial sums on the host in
// shard order.  Philox counters carry the GLOBAL replica number, so a trajectory does not depend on the number of
// shards (tests/test_gpu_api.py::test_fleet_*).  Host code only; included at the end of kmos_b200.cu.
#pragma once

struct kmos_b200_fleet {
    kmos_b200_model* model = nullptr;
    int32_t R = 0;
    std::vector<kmos_b200_batch*> shard;  // non-empty shards only
    std::vector<int32_t> lo, n;           // first global replica and replica count of each
};

#define KB_FLEET_EACH(f, k, expr)                                  \
    do {                                                           \
        for (size_t k = 0; k < (f)->shard.size(); ++k) {           \
            const int rc_ = (expr);                                \
            if (rc_ != KMOS_B200_OK) return rc_;                   \
        }                                                          \
        return KMOS_B200_OK;                                       \
    } while (0)

extern "C" void kmos_b200_fleet_destroy(kmos_b200_fleet* f) {
    if (!f) return;
    for (kmos_b200_batch* b : f->shard) kmos_b200_batch_destroy(b);
    delete f;
}

extern "C" int kmos_b200_fleet_create(kmos_b200_model* m, int32_t n_replicas, const int32_t size[3], const uint64_t* seeds,
                                      const int32_t* gpu_ids, int32_t n_gpus, kmos_b200_fleet** out) {
    if (!m || !out || !gpu_ids || n_gpus < 1 || n_replicas < 1)
        return set_err(KMOS_B200_ERR_ARG, "fleet_create: model, gpu_ids[n_gpus >= 1] and n_replicas >= 1 required");
    kmos_b200_fleet* f = new kmos_b200_fleet;
    f->model = m;
    f->R = n_replicas;
    for (int32_t k = 0; k < n_gpus; ++k) {
        const int32_t lo = (int32_t)((int64_t)n_replicas * k / n_gpus), hi = (int32_t)((int64_t)n_replicas * (k + 1) / n_gpus);
        if (hi == lo) continue;  // fewer replicas than GPUs: this one stays idle
        kmos_b200_batch* b = nullptr;
        int rc = kmos_b200_batch_create(m, hi - lo, size, gpu_ids[k], &b);
        if (rc == KMOS_B200_OK) {
            f->shard.push_back(b);
            f->lo.push_back(lo);
            f->n.push_back(hi - lo);
            std::vector<uint32_t> ids((size_t)(hi - lo));
            std::vector<uint64_t> key((size_t)(hi - lo));
            for (int32_t r = lo; r < hi; ++r) {
                ids[r - lo] = (uint32_t)r;
                key[r - lo] = seeds ? seeds[r] : (uint64_t)r;
            }
            rc = kmos_b200_set_seeds(b, key.data(), ids.data());
        }
        if (rc != KMOS_B200_OK) {
            const std::string why = g_err;  // batch_destroy may overwrite it
            kmos_b200_fleet_destroy(f);
            return set_err(rc, "fleet_create, shard on device " + std::to_string(gpu_ids[k]) + ": " + why);
        }
    }
    *out = f;
    return KMOS_B200_OK;
}

extern "C" int kmos_b200_fleet_n_shards(const kmos_b200_fleet* f) { return (int)f->shard.size(); }

extern "C" kmos_b200_batch* kmos_b200_fleet_shard(kmos_b200_fleet* f, int32_t k, int32_t* first_replica, int32_t* n_replicas) {
    if (!f || k < 0 || k >= (int32_t)f->shard.size()) return nullptr;
    if (first_replica) *first_replica = f->lo[k];
    if (n_replicas) *n_replicas = f->n[k];
    return f->shard[k];
}

extern "C" int kmos_b200_fleet_attach_proclist(kmos_b200_fleet* f, const char* so_path) {
    KB_FLEET_EACH(f, k, kmos_b200_batch_attach_proclist(f->shard[k], so_path));
}
extern "C" int kmos_b200_fleet_select_kernel(kmos_b200_fleet* f, int32_t kind) {
    KB_FLEET_EACH(f, k, kmos_b200_select_kernel(f->shard[k], kind));
}
extern "C" int kmos_b200_fleet_set_rates(kmos_b200_fleet* f, const double* rates) {
    const size_t P = (size_t)f->model->h.n_proc;
    KB_FLEET_EACH(f, k, kmos_b200_set_rates(f->shard[k], rates + (size_t)f->lo[k] * P));
}
extern "C" int kmos_b200_fleet_set_otf_lut(kmos_b200_fleet* f, const double* lut) {
    const size_t W = (size_t)f->model->h.lut_total;
    KB_FLEET_EACH(f, k, kmos_b200_set_otf_lut(f->shard[k], lut + (size_t)f->lo[k] * W));
}
extern "C" int kmos_b200_fleet_init_state(kmos_b200_fleet* f, int32_t layer) {
    KB_FLEET_EACH(f, k, kmos_b200_init_state(f->shard[k], layer));
}
// enqueue on every shard's stream first: the GPUs step concurrently, nothing here waits
extern "C" int kmos_b200_fleet_do_kmc_steps(kmos_b200_fleet* f, int64_t n) {
    KB_FLEET_EACH(f, k, kmos_b200_do_kmc_steps(f->shard[k], n));
}
extern "C" int kmos_b200_fleet_synchronize(kmos_b200_fleet* f) {
    KB_FLEET_EACH(f, k, kmos_b200_synchronize(f->shard[k]));
}

// getters: shard k fills out + lo[k] * (elements per replica)
extern "C" int kmos_b200_fleet_get_kmc_time(kmos_b200_fleet* f, double* out) {
    KB_FLEET_EACH(f, k, kmos_b200_get_kmc_time(f->shard[k], out + f->lo[k]));
}
extern "C" int kmos_b200_fleet_get_kmc_step(kmos_b200_fleet* f, int64_t* out) {
    KB_FLEET_EACH(f, k, kmos_b200_get_kmc_step(f->shard[k], out + f->lo[k]));
}
extern "C" int kmos_b200_fleet_get_status(kmos_b200_fleet* f, int32_t* out) {
    KB_FLEET_EACH(f, k, kmos_b200_get_status(f->shard[k], out + f->lo[k]));
}
extern "C" int kmos_b200_fleet_get_procstat(kmos_b200_fleet* f, int64_t* out) {
    const size_t P = (size_t)f->model->h.n_proc;
    KB_FLEET_EACH(f, k, kmos_b200_get_procstat(f->shard[k], out + (size_t)f->lo[k] * P));
}
extern "C" int kmos_b200_fleet_get_integ_rates(kmos_b200_fleet* f, double* out) {
    const size_t P = (size_t)f->model->h.n_proc;
    KB_FLEET_EACH(f, k, kmos_b200_get_integ_rates(f->shard[k], out + (size_t)f->lo[k] * P));
}
extern "C" int kmos_b200_fleet_get_nr_of_sites(kmos_b200_fleet* f, int32_t* out) {
    const size_t P = (size_t)f->model->h.n_proc;
    KB_FLEET_EACH(f, k, kmos_b200_get_nr_of_sites(f->shard[k], out + (size_t)f->lo[k] * P));
}
extern "C" int kmos_b200_fleet_get_lattice(kmos_b200_fleet* f, int32_t* out) {
    if (f->shard.empty()) return KMOS_B200_OK;
    const size_t V = (size_t)kmos_b200_batch_volume(f->shard[0]);
    KB_FLEET_EACH(f, k, kmos_b200_get_lattice(f->shard[k], out + (size_t)f->lo[k] * V));
}
extern "C" int kmos_b200_fleet_get_occupation(kmos_b200_fleet* f, double* out) {
    const size_t W = (size_t)f->model->h.n_species * f->model->h.spuck;
    KB_FLEET_EACH(f, k, kmos_b200_get_occupation(f->shard[k], out + (size_t)f->lo[k] * W));
}

// The tally of kmos_b200_reduce_tallies over all shards: every GPU reduces its replicas on the device, the
// n_groups x tally_words partial sums are added here in shard order (what the NCCL all-reduce does between ranks).
extern "C" int kmos_b200_fleet_reduce_tallies(kmos_b200_fleet* f, const int32_t* group_of, int32_t n_groups, double* host_out) {
    if (!host_out || n_groups < 1) return set_err(KMOS_B200_ERR_ARG, "fleet_reduce_tallies: host_out and n_groups >= 1 required");
    if (f->shard.empty()) return KMOS_B200_OK;
    const size_t words = (size_t)n_groups * kmos_b200_tally_words(f->shard[0]);
    std::vector<double> part(words);
    std::fill(host_out, host_out + words, 0.0);
    for (size_t k = 0; k < f->shard.size(); ++k) {
        const int rc = kmos_b200_reduce_tallies(f->shard[k], group_of ? group_of + f->lo[k] : nullptr, n_groups, nullptr, part.data());
        if (rc != KMOS_B200_OK) return rc;
        for (size_t i = 0; i < words; ++i) host_out[i] += part[i];
    }
    return KMOS_B200_OK;
}
