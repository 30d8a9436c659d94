// kb_interp.h -- generic (any backend, any lattice size) kMC engine: one thread steps one replica and
// interprets the model's statement byte-code against state that lives in global memory.
//
// It is the engine for initialize_state / touchup (which define the initial avail_sites order), for the
// lat_int and otf backends and for lattices too large for shared memory; the shared-memory warp kernel
// (kb_smem.cuh) replaces it for local_smart models that fit.
//
// Restated reference routines (file:line in the kmos checkout):
//   do_kmc_steps                  kmos/fortran_src/proclist_generic_subroutines.mpy:1-44
//   initialize_state              kmos/fortran_src/proclist_generic_subroutines.mpy:236-304
//   update_accum_rate             base.mpy:603-623        otf: base_otf.f90:687-717
//   update_clocks                 base.mpy:1123-1161
//   update_integ_rate             base.mpy:626-645        otf: base_otf.f90:719-739
//   determine_procsite            base.mpy:1075-1120      otf: base_otf.f90:1213-1277
//   interval_search_real          base.mpy:1234-1338
//   add_proc / del_proc / can_do  base.mpy:211-321        otf: base_otf.f90:221-364
//   replace_species               base.mpy:1187-1231
//   run_proc_nr and callees       generated; byte-code from kmos_b200/tables.py
#pragma once
#include <math.h>

#include "kb_common.h"

#define KB_CALL_DEPTH 6

template <typename idx_t>
struct KbInterp {
    const KbModelView& m;
    const KbGeom& g;
    KbReplica<idx_t>& r;
    int nr_vars[KB_MAX_VARS];

    KB_HDN KbInterp(const KbModelView& m_, const KbGeom& g_, KbReplica<idx_t>& r_) : m(m_), g(g_), r(r_) {}

    // ---- base ----------------------------------------------------------------------------------
    KB_HD int anchor_n(int proc) const { return m.procsite[proc - 1]; }  // converted to a site type at load

    // avail_sites(proc, site, 2) /= 0 ; a process is never registered on another site type
    KB_HD int pos_of(int proc, int cell, int n) const {
        if (n != anchor_n(proc)) return 0;
        return (int)r.p2[(size_t)(proc - 1) * g.ncells + cell];
    }

    KB_HD void del_proc(int proc, int cell, int n) {
        if (proc <= 0) return;  // base_lat_int.mpy:249
        size_t row = (size_t)(proc - 1) * g.ncells;
        int nq = r.nsites[proc - 1];
        int pos = (int)r.p2[row + cell];
        if (n != anchor_n(proc) || pos == 0) { fail(KB_BAD_MODEL); return; }
        if (pos < nq) {
            idx_t last = r.p1[row + nq - 1];
            r.p1[row + pos - 1] = last;
            r.p1[row + nq - 1] = 0;
            if (m.backend == KB_BACKEND_OTF) {
                double* rm = r.rates_matrix + (size_t)(proc - 1) * (g.ncells + 1);
                if (r.blk) {  // the moved rate changes block, the freed one leaves
                    double* bl = r.blk + (size_t)(proc - 1) * r.blk_n;
                    bl[(pos - 1) >> 8] += rm[nq - 1] - rm[pos - 1];
                    bl[(nq - 1) >> 8] -= rm[nq - 1];
                }
                rm[g.ncells] = KB_SUB(rm[g.ncells], rm[pos - 1]);
                rm[pos - 1] = rm[nq - 1];
                rm[nq - 1] = 0.0;
            }
            r.p2[row + last] = (idx_t)pos;
        } else {
            r.p1[row + pos - 1] = 0;
            if (m.backend == KB_BACKEND_OTF) {
                double* rm = r.rates_matrix + (size_t)(proc - 1) * (g.ncells + 1);
                if (r.blk) r.blk[(size_t)(proc - 1) * r.blk_n + ((pos - 1) >> 8)] -= rm[pos - 1];
                rm[g.ncells] = KB_SUB(rm[g.ncells], rm[pos - 1]);
                rm[pos - 1] = 0.0;
            }
        }
        r.p2[row + cell] = 0;
        r.nsites[proc - 1] = nq - 1;
    }

    KB_HD void add_proc(int proc, int cell, int n, double rate) {
        if (proc <= 0) return;  // base_lat_int.mpy:299
        if (n != anchor_n(proc)) { fail(KB_BAD_MODEL); return; }
        size_t row = (size_t)(proc - 1) * g.ncells;
        int nq = r.nsites[proc - 1] + 1;
        if (nq > g.ncells || r.p2[row + cell] != 0) { fail(KB_CAPACITY); return; }
        r.nsites[proc - 1] = nq;
        r.p1[row + nq - 1] = (idx_t)cell;
        r.p2[row + cell] = (idx_t)nq;
        if (m.backend == KB_BACKEND_OTF) {
            double* rm = r.rates_matrix + (size_t)(proc - 1) * (g.ncells + 1);
            if (r.blk) r.blk[(size_t)(proc - 1) * r.blk_n + ((nq - 1) >> 8)] += rate;
            rm[g.ncells] = KB_ADD(rm[g.ncells], rate);
            rm[nq - 1] = rate;
        }
    }

    KB_HD void update_rates_matrix(int proc, int cell, double rate) {
        double* rm = r.rates_matrix + (size_t)(proc - 1) * (g.ncells + 1);
        int pos = (int)r.p2[(size_t)(proc - 1) * g.ncells + cell];
        if (r.blk) r.blk[(size_t)(proc - 1) * r.blk_n + ((pos - 1) >> 8)] += rate - rm[pos - 1];
        rm[g.ncells] = KB_SUB(KB_ADD(rm[g.ncells], rate), rm[pos - 1]);
        rm[pos - 1] = rate;
    }

    KB_HD void fail(int code) {
        if (r.status == KB_OK) r.status = code;
    }

    KB_HD void replace_species(int cell, int n, int old_species, int new_species) {
        int idx = cell * m.spuck + n - 1;
        int found = r.lattice[idx] == KB_NULL_SPECIES ? -1 : (int)r.lattice[idx];
        if (found != old_species) {
            if (r.status == KB_OK) {
                r.status = KB_SPECIES_MISMATCH;
                r.err[0] = old_species; r.err[1] = new_species; r.err[2] = found;
                r.err[3] = idx + 1; r.err[4] = (int32_t)r.kmc_step;
            }
            return;
        }
        r.lattice[idx] = new_species < 0 ? KB_NULL_SPECIES : (uint8_t)new_species;
    }

    // ---- byte-code ------------------------------------------------------------------------------
    struct Site { int cell, n; };
    KB_HD Site site_of(const int base[4], const int32_t* off) const {
        Site s;
        s.cell = kb_cell_of(m, g, base[0] + off[0], base[1] + off[1], base[2] + off[2]);
        s.n = base[3] + off[3];
        return s;
    }
    KB_HD int species_at(const int base[4], const int32_t* off) const {
        Site s = site_of(base, off);
        uint8_t v = r.lattice[s.cell * m.spuck + s.n - 1];
        return v == KB_NULL_SPECIES ? -1 : (int)v;
    }

    // nli_<group>(cell) / gr_<proc>(cell): flat walk, SELECT / RETURN / INC only
    KB_HD int eval_func(int rid, const int cell[4]) {
        const int32_t* pc = m.code + m.routines[2 * rid];
        const int32_t* end = pc + m.routines[2 * rid + 1];
        while (pc < end) {
            switch (*pc) {
            case KB_OP_SELECT: pc = select_target(pc, cell); break;
            case KB_OP_JUMP: pc += 2 + pc[1]; break;
            case KB_OP_RETURN: return pc[1];
            case KB_OP_INC: nr_vars[pc[1]]++; pc += 2; break;
            default: fail(KB_BAD_MODEL); return 0;
            }
        }
        return 0;
    }

    // OP_SELECT dx dy dz dn ncases total | ncases x (OP_CASE mask len body... OP_JUMP rest)
    KB_HD const int32_t* select_target(const int32_t* pc, const int base[4]) const {
        int species = species_at(base, pc + 1);
        int ncases = pc[5];
        const int32_t* c = pc + 7;
        for (int i = 0; i < ncases; ++i) {
            int32_t mask = c[1];
            if (mask == -1 || (species >= 0 && ((mask >> species) & 1))) return c + 3;
            c += 3 + c[2];
        }
        return pc + 7 + pc[6];
    }

    KB_HD double eval_gr(int gid, const int base[4], const int32_t* off) {
        const int32_t* gdesc = m.gr + (size_t)gid * KB_GR_STRIDE;
        int cell[4] = {base[0] + off[0], base[1] + off[1], base[2] + off[2], base[3] + off[3]};
        for (int k = 0; k < KB_MAX_VARS; ++k) nr_vars[k] = 0;
        eval_func(gdesc[0], cell);
        int idx = 0, stride = 1;
        for (int k = 0; k < gdesc[2]; ++k) {
            idx += nr_vars[k] * stride;
            stride *= gdesc[4 + k];
        }
        return r.lut[gdesc[3] + idx];
    }

    // run one routine (and the routines it calls) on base coordinate base0
    KB_HDN void exec(int rid, const int base0[4]) {
        const int32_t* ret_pc[KB_CALL_DEPTH];
        const int32_t* ret_end[KB_CALL_DEPTH];
        int ret_base[KB_CALL_DEPTH][4];
        int depth = 0;
        int base[4] = {base0[0], base0[1], base0[2], base0[3]};
        const int32_t* pc = m.code + m.routines[2 * rid];
        const int32_t* end = pc + m.routines[2 * rid + 1];
        for (;;) {
            if (pc >= end) {
                if (depth == 0) return;
                --depth;
                pc = ret_pc[depth]; end = ret_end[depth];
                for (int i = 0; i < 4; ++i) base[i] = ret_base[depth][i];
                continue;
            }
            switch (*pc) {
            case KB_OP_REPLACE: {
                Site s = site_of(base, pc + 1);
                replace_species(s.cell, s.n, pc[5], pc[6]);
                pc += 7;
                break;
            }
            case KB_OP_IF_CAN: {
                Site s = site_of(base, pc + 2);
                if (pos_of(pc[1], s.cell, s.n)) pc += 7; else pc += 7 + pc[6];
                break;
            }
            case KB_OP_DEL: {
                Site s = site_of(base, pc + 2);
                del_proc(pc[1], s.cell, s.n);
                pc += 6;
                break;
            }
            case KB_OP_ADD: {
                Site s = site_of(base, pc + 2);
                add_proc(pc[1], s.cell, s.n, 0.0);
                pc += 6;
                break;
            }
            case KB_OP_DEL_NLI:
            case KB_OP_ADD_NLI: {
                int cell[4] = {base[0] + pc[2], base[1] + pc[3], base[2] + pc[4], base[3] + pc[5]};
                int proc = eval_func(pc[1], cell);
                Site s = site_of(base, pc + 6);
                if (*pc == KB_OP_DEL_NLI) del_proc(proc, s.cell, s.n); else add_proc(proc, s.cell, s.n, 0.0);
                pc += 10;
                break;
            }
            case KB_OP_ADD_RATE: {
                Site s = site_of(base, pc + 2);
                add_proc(pc[1], s.cell, s.n, eval_gr(pc[6], base, pc + 7));
                pc += 11;
                break;
            }
            case KB_OP_UPD_RATE: {
                Site s = site_of(base, pc + 2);
                update_rates_matrix(pc[1], s.cell, eval_gr(pc[6], base, pc + 7));
                pc += 11;
                break;
            }
            case KB_OP_SELECT: pc = select_target(pc, base); break;
            case KB_OP_JUMP: pc += 2 + pc[1]; break;
            case KB_OP_DEL_ALL: {
                Site s = site_of(base, pc + 1);
                for (int p = 1; p <= m.n_proc; ++p)
                    if (pos_of(p, s.cell, s.n)) del_proc(p, s.cell, s.n);
                pc += 5;
                break;
            }
            case KB_OP_CALL: {
                if (depth == KB_CALL_DEPTH) { fail(KB_BAD_MODEL); return; }
                ret_pc[depth] = pc + 6; ret_end[depth] = end;
                for (int i = 0; i < 4; ++i) ret_base[depth][i] = base[i];
                ++depth;
                for (int i = 0; i < 4; ++i) base[i] += pc[2 + i];
                int callee = pc[1];
                pc = m.code + m.routines[2 * callee];
                end = pc + m.routines[2 * callee + 1];
                break;
            }
            default: fail(KB_BAD_MODEL); return;
            }
            if (r.status == KB_BAD_MODEL) return;
        }
    }

    // ---- step loop --------------------------------------------------------------------------------
    KB_HD void update_accum_rate() {
        int P = m.n_proc;
        if (m.backend == KB_BACKEND_OTF) {
            double acc = 0.0;
            for (int i = 0; i < P; ++i) {
                double* rm = r.rates_matrix + (size_t)i * (g.ncells + 1);
                double tot = 0.0;
                int nq = r.nsites[i];
                for (int j = 0; j < nq; ++j) tot = KB_ADD(tot, rm[j]);
                rm[g.ncells] = tot;
                acc = (i == 0) ? tot : KB_ADD(acc, tot);
                r.accum[i] = acc;
            }
            return;
        }
        double acc = KB_MUL((double)r.nsites[0], r.rates[0]);
        r.accum[0] = acc;
        for (int i = 1; i < P; ++i) {
            acc = KB_ADD(acc, KB_MUL((double)r.nsites[i], r.rates[i]));
            r.accum[i] = acc;
        }
    }

    KB_HD void update_integ_rate() {
        for (int i = 0; i < m.n_proc; ++i) {
            double w = (m.backend == KB_BACKEND_OTF) ? r.rates_matrix[(size_t)i * (g.ncells + 1) + g.ncells]
                                                     : KB_MUL((double)r.nsites[i], r.rates[i]);
            r.integ[i] = KB_ADD(r.integ[i], KB_MUL(w, r.kmc_time_step));
        }
    }

    // returns 1-based index, 0 = nothing available (the reference prints a dead-lock message and stops)
    static KB_HD int interval_search_real(const double* arr, int size, double value) {
        int left = 1, right = size, mid;
        for (;;) {
            mid = (right + left) >> 1;
            if (left >= right) break;
            if (value < arr[mid - 1]) right = mid; else left = mid + 1;
        }
        if (arr[mid - 1] == 0.) {
            for (;;) {
                if (mid > size) return 0;
                if (arr[mid - 1] > 0.) {
                    if (mid >= size) return 0;
                    break;
                }
                mid = mid + 1;
            }
        }
        for (;;) {
            if (mid == 1) break;
            if (arr[mid - 2] >= arr[mid - 1]) mid = mid - 1; else break;
        }
        return mid;
    }

    KB_HD bool determine_procsite(double ran_proc, double ran_site, int* proc, int* cell) {
        int P = m.n_proc;
        int p = interval_search_real(r.accum, P, KB_MUL(ran_proc, r.accum[P - 1]));
        if (p == 0 || r.nsites[p - 1] <= 0) { fail(KB_DEADLOCK); return false; }
        int nq = r.nsites[p - 1];
        int k;
        if (m.backend == KB_BACKEND_OTF) {
            const double* rm = r.rates_matrix + (size_t)(p - 1) * (g.ncells + 1);
            double acc = rm[0];
            r.accum_proc[0] = acc;
            for (int i = 1; i < nq; ++i) {
                acc = KB_ADD(acc, rm[i]);
                r.accum_proc[i] = acc;
            }
            k = interval_search_real(r.accum_proc, nq, KB_MUL(ran_site, r.accum_proc[nq - 1]));
            if (k == 0) { fail(KB_DEADLOCK); return false; }
        } else {
            k = (int)KB_ADD(1.0, KB_MUL(ran_site, (double)nq));  // int(1+ran_site*nr_of_sites(proc))
            if (k > nq) k = nq;
        }
        *proc = p;
        *cell = (int)r.p1[(size_t)(p - 1) * g.ncells + k - 1];
        return true;
    }

    KB_HD void cell_coords(int cell, int base[4]) const {
        base[0] = cell % g.size[0];
        int c = cell / g.size[0];
        base[1] = c % g.size[1];
        base[2] = c / g.size[1];
    }

    KB_HD void run_proc_nr(int proc, int cell) {
        int lsite[4];
        r.procstat[proc - 1]++;
        cell_coords(cell, lsite);
        lsite[3] = anchor_n(proc);
        exec(m.runproc[proc - 1], lsite);
    }

    KB_HDN void do_kmc_steps(int64_t n) {
        for (int64_t i = 0; i < n && r.status == KB_OK; ++i) {
            double ran_time, ran_proc, ran_site;
            kb_philox_step(r.seed, r.replica, (uint64_t)r.kmc_step, &ran_time, &ran_proc, &ran_site);
            update_accum_rate();
            double total = r.accum[m.n_proc - 1];
            if (!(total > 0.)) { fail(KB_DEADLOCK); break; }
            r.kmc_time_step = -log(ran_time) / total;  // update_clocks
            r.kmc_time = KB_ADD(r.kmc_time, r.kmc_time_step);
            r.kmc_step = r.kmc_step + 1;
            update_integ_rate();
            int proc, cell;
            if (!determine_procsite(ran_proc, ran_site, &proc, &cell)) break;
            run_proc_nr(proc, cell);
        }
    }

    // initialize_state: default species everywhere, then touchup cell by cell (z, y, x-fastest order)
    KB_HDN void init_state(int layer) {
        const uint8_t null_fill = m.null_species < 0 ? (uint8_t)KB_NULL_SPECIES : (uint8_t)m.null_species;
        for (int i = 0; i < g.volume; ++i) r.lattice[i] = null_fill;
        for (int q = 0; q < m.n_proc; ++q) {
            r.nsites[q] = 0; r.integ[q] = 0.0; r.accum[q] = 0.0; r.procstat[q] = 0;
        }
        for (size_t i = 0; i < (size_t)m.n_proc * g.ncells; ++i) { r.p1[i] = 0; r.p2[i] = 0; }
        if (m.backend == KB_BACKEND_OTF)
            for (size_t i = 0; i < (size_t)m.n_proc * (g.ncells + 1); ++i) r.rates_matrix[i] = 0.0;
        r.kmc_time = 0.0; r.kmc_time_step = 0.0; r.kmc_step = 0; r.status = KB_OK;
        for (int pass = 0; pass < 2; ++pass)
            for (int k = 0; k < g.size[2]; ++k)
                for (int j = 0; j < g.size[1]; ++j)
                    for (int i = 0; i < g.size[0]; ++i) {
                        int base[4] = {i, j, k, 0};
                        exec(m.init[2 * layer + pass], base);
                    }
    }

    // proclist.recalculate_rates_matrix (proclist_generic_subroutines.mpy:307-325) + base.reaccumulate_rates_matrix
    // (base_otf.f90:366-387): every registered (process, site) gets the current gr_<proc> value, then every
    // row total is re-added from scratch in memory-address order (which makes the visiting order irrelevant)
    KB_HDN void recalculate_rates_matrix() {
        if (m.backend != KB_BACKEND_OTF) return;
        const int32_t zero_off[4] = {0, 0, 0, 0};
        for (int gid = 0; gid < m.n_gr; ++gid) {
            const int proc = m.gr[(size_t)gid * KB_GR_STRIDE + 1];
            const size_t row = (size_t)(proc - 1) * g.ncells;
            double* rm = r.rates_matrix + (size_t)(proc - 1) * (g.ncells + 1);
            const int nq = r.nsites[proc - 1];
            for (int pos = 0; pos < nq; ++pos) {
                int cell = (int)r.p1[row + pos];
                int base[4];
                base[0] = cell % g.size[0]; cell /= g.size[0];
                base[1] = cell % g.size[1]; cell /= g.size[1];
                base[2] = cell; base[3] = 0;
                rm[pos] = eval_gr(gid, base, zero_off);
            }
        }
        for (int proc = 1; proc <= m.n_proc; ++proc) {
            double* rm = r.rates_matrix + (size_t)(proc - 1) * (g.ncells + 1);
            const int nq = r.nsites[proc - 1];
            double tot = 0.0;
            for (int pos = 0; pos < nq; ++pos) tot = KB_ADD(tot, rm[pos]);
            rm[g.ncells] = tot;
        }
        update_accum_rate();
    }

    // KMC_Model._set_configuration + _adjust_database (kmos/run/__init__.py:1411-1457): the lattice has
    // been overwritten by the caller; touch up x-outermost, without clearing avail_sites first.
    KB_HDN void adjust_database(int layer) {
        for (int i = 0; i < g.size[0]; ++i)
            for (int j = 0; j < g.size[1]; ++j)
                for (int k = 0; k < g.size[2]; ++k) {
                    int base[4] = {i, j, k, 0};
                    exec(m.init[2 * layer + 1], base);
                }
        update_accum_rate();
    }
};
