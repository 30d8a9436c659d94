"""lat_int / otf semantics pinned by a parser-independent execution of the reference's committed Fortran.

The reference's golden trajectory pins local_smart only (its lat_int/otf logs are byte-copies of the
local_smart one).  Here the generated Fortran the reference keeps under version control
for its export tests is executed statement by statement by tests/fortran_exec.py -- which shares no code with
kmos_b200/fortran_ir.py / tables.py -- and compared with the oracle (which runs the byte-code those two files
produce from the very same text) over 10 000 events under the shared Philox stream: lattice, nr_of_sites,
both planes of avail_sites, procstat bit-exact, kmc_time <= 1e-12, and for otf every row of rates_matrix.

Needs /root/reference (this container); skipped where it is absent (the GPU box).
"""
import os

import numpy as np
import pytest

from kmos_b200 import fortran_ir, otf as otf_mod, tables
from oracle import oracle

import fortran_exec

REF = "/root/reference/tests/export_test"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")

CASES = [
    # directory, backend, lattice, events, checkpoint interval
    ("reference_export", "local_smart", [6, 5], 10000, 1000),          # the pinned backend, as a control
    ("reference_export_lat_int", "lat_int", [6, 5], 10000, 1000),      # RuO2: 36 nli_* trees, 36 run_proc_*
    ("reference_export_otf", "otf", [5, 4], 10000, 1000),              # RuO2 as otf: gr_/rate_ without bystanders
    ("reference_export_intZGB_otf", "otf", [7, 6], 10000, 1000),       # bystander-dependent rates (nr_vars)
    ("reference_pdopd_lat_int", "lat_int", [4, 4], 6000, 1000),        # Pd(100)/PdO: two lattices, null species
    ("reference_pdopd_local_smart", "local_smart", [4, 4], 6000, 1000),
]


@pytest.mark.parametrize("dirname,backend,size,n_events,every", CASES)
def test_fortran_text_executed_directly_matches_the_oracle(dirname, backend, size, n_events, every):
    path = os.path.join(REF, dirname)
    fm = fortran_exec.FortranModel(path)
    assert fm.backend == backend
    ir = fortran_ir.parse_export_dir(path, backend)
    P = len(ir["procs"])
    assert fm.nr_of_proc == P and [fm.proc_names[i + 1] for i in range(P)] == [p.lower() for p in ir["procs"]]
    rng = np.random.RandomState(len(dirname))
    rates = np.exp(rng.uniform(-1.0, 1.0, P))
    userpar, lut = None, None
    if backend == "otf":
        names = fm.userpar_names()
        assert [n.lower() for n in ir.get("userpar", [])] == names
        values = np.exp(rng.uniform(-0.4, 0.4, len(names)))
        ir["parameters"] = {n: {"value": repr(float(v))} for n, v in zip(ir.get("userpar", []), values)}
        userpar = [float(v) for v in values]
        assert not ir.get("chempots")
    blob, info = tables.build_blob(ir)
    if backend == "otf":
        lut = otf_mod.build_lut(ir, info, rates)
    seed, replica = 4242, 7
    o = oracle.Oracle(blob, size, seed=seed, replica=replica, rates=rates, lut=lut)
    ex = fortran_exec.Executor(fm, size, rates, seed, replica, oracle.philox_step, userpar=userpar,
                               layer=ir["layers"][ir["default_layer"]].lower() if "layers" in ir else None)

    def compare(tag):
        assert np.array_equal(np.asarray(ex.lattice), o.lattice), "lattice differs %s" % tag
        assert np.array_equal(np.asarray(ex.nr_of_sites[1:]), o.nr_of_sites), "nr_of_sites differs %s" % tag
        assert np.array_equal(ex.avail_sites_array(), o.avail_sites), "avail_sites differ %s" % tag
        assert np.array_equal(np.asarray(ex.procstat[1:]), o.procstat), "procstat differs %s" % tag
        if o.kmc_time > 0:
            assert abs(ex.kmc_time - o.kmc_time) <= 1e-12 * o.kmc_time, tag
        if backend == "otf":
            for p in range(1, P + 1):
                row = o.rates_matrix_row(p)
                n = ex.nr_of_sites[p]
                np.testing.assert_allclose(np.asarray(ex.rates_matrix[p][1:n + 1]), row[:n], rtol=0, atol=0,
                                           err_msg="rates_matrix row %d %s" % (p, tag))

    compare("after initialize_state")
    done = 0
    while done < n_events:
        for _ in range(every):
            ex.step()
        assert o.do_steps(every) == 0
        done += every
        compare("after %d events" % done)
    assert ex.kmc_step == o.kmc_step == n_events
    assert sum(ex.procstat) == n_events
