// Host build of kmos_b200/csrc/kb_interp.h (the generic CUDA engine's source) for CPU-side validation
// against the oracle.  TEST INFRASTRUCTURE: compiled by tests/test_interp_host.py with
//   g++ -O2 -ffp-contract=off -shared -fPIC -I kmos_b200/csrc
// The product path only ever runs this header as device code.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kb_interp.h"

struct Harness {
    std::vector<int32_t> blob;
    KbModelView m;
    KbGeom g;
    KbReplica<uint16_t> r;
    std::vector<uint8_t> lattice;
    std::vector<int32_t> nsites;
    std::vector<uint16_t> p1, p2;
    std::vector<double> rates, integ, accum, rates_matrix, accum_proc, lut;
    std::vector<int64_t> procstat;
};

extern "C" {

// KbInterp::interval_search_real, the routine the CUDA kernels compile (unit tests)
int kbh_interval_search_real(const double* arr, int size, double value) {
    return KbInterp<uint16_t>::interval_search_real(arr, size, value);
}

Harness* kbh_create(const int32_t* blob, int64_t n, const int32_t size[3], uint64_t seed, uint32_t replica) {
    Harness* h = new Harness;
    h->blob.assign(blob, blob + n);
    if (!kb_model_view(h->blob.data(), n, h->blob.data(), &h->m)) { delete h; return nullptr; }
    for (int i = 0; i < 3; ++i) h->g.size[i] = i < h->m.dim ? size[i] : 1;
    h->g.ncells = h->g.size[0] * h->g.size[1] * h->g.size[2];
    h->g.volume = h->g.ncells * h->m.spuck;
    int P = h->m.n_proc, C = h->g.ncells;
    h->lattice.assign(h->g.volume, KB_NULL_SPECIES);
    h->nsites.assign(P, 0); h->p1.assign((size_t)P * C, 0); h->p2.assign((size_t)P * C, 0);
    h->rates.assign(P, 0); h->integ.assign(P, 0); h->accum.assign(P, 0); h->procstat.assign(P, 0);
    h->rates_matrix.assign((size_t)P * (C + 1), 0); h->accum_proc.assign(C, 0);
    h->lut.assign(h->m.lut_total > 0 ? h->m.lut_total : 1, 0);
    memset(&h->r, 0, sizeof h->r);
    h->r.lattice = h->lattice.data(); h->r.nsites = h->nsites.data(); h->r.p1 = h->p1.data(); h->r.p2 = h->p2.data();
    h->r.rates = h->rates.data(); h->r.integ = h->integ.data(); h->r.accum = h->accum.data();
    h->r.procstat = h->procstat.data(); h->r.rates_matrix = h->rates_matrix.data();
    h->r.accum_proc = h->accum_proc.data(); h->r.lut = h->lut.data();
    h->r.seed = seed; h->r.replica = replica;
    return h;
}
void kbh_destroy(Harness* h) { delete h; }
void kbh_set_rates(Harness* h, const double* r) { memcpy(h->rates.data(), r, h->rates.size() * 8); }
void kbh_set_lut(Harness* h, const double* l) { memcpy(h->lut.data(), l, (size_t)h->m.lut_total * 8); }
int kbh_init_state(Harness* h, int layer) {
    KbInterp<uint16_t> it(h->m, h->g, h->r);
    it.init_state(layer);
    return h->r.status;
}
int kbh_set_configuration(Harness* h, const int32_t* species, int layer) {
    for (int i = 0; i < h->g.volume; ++i) h->lattice[i] = species[i] < 0 ? KB_NULL_SPECIES : (uint8_t)species[i];
    KbInterp<uint16_t> it(h->m, h->g, h->r);
    it.adjust_database(layer);
    return h->r.status;
}
int kbh_do_steps(Harness* h, int64_t n) {
    KbInterp<uint16_t> it(h->m, h->g, h->r);
    it.do_kmc_steps(n);
    return h->r.status;
}
double kbh_kmc_time(Harness* h) { return h->r.kmc_time; }
int64_t kbh_kmc_step(Harness* h) { return h->r.kmc_step; }
void kbh_get_lattice(Harness* h, int32_t* out) {
    for (int i = 0; i < h->g.volume; ++i) out[i] = h->lattice[i] == KB_NULL_SPECIES ? -1 : h->lattice[i];
}
void kbh_get_procstat(Harness* h, int64_t* out) { memcpy(out, h->procstat.data(), h->procstat.size() * 8); }
void kbh_get_nsites(Harness* h, int32_t* out) { memcpy(out, h->nsites.data(), h->nsites.size() * 4); }
void kbh_get_integ(Harness* h, double* out) { memcpy(out, h->integ.data(), h->integ.size() * 8); }
// reference layout avail_sites[P][volume][2], 1-based contents
void kbh_get_avail(Harness* h, int32_t* out) {
    int P = h->m.n_proc, C = h->g.ncells, V = h->g.volume, sp = h->m.spuck;
    memset(out, 0, (size_t)P * V * 2 * 4);
    for (int q = 0; q < P; ++q) {
        int n = h->m.procsite[q];
        for (int k = 0; k < h->nsites[q]; ++k) out[((size_t)q * V + k) * 2] = h->p1[(size_t)q * C + k] * sp + n;
        for (int c = 0; c < C; ++c) out[((size_t)q * V + c * sp + n - 1) * 2 + 1] = h->p2[(size_t)q * C + c];
    }
}
}
