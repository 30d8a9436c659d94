"""CPU-side validation of the generic CUDA engine's SOURCE (kmos_b200/csrc/kb_interp.h compiled for the
host) against the oracle: same Philox stream -> bit-identical lattice, procstat, nr_of_sites and
avail_sites (both planes), kmc_time within 1e-12.  Not a product path -- the product runs it on the GPU
(tests/test_gpu_parity.py does the same comparison through the C-ABI)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO, load_model
from kmos_b200 import otf as otf_mod
from oracle import oracle

SO = os.path.join(REPO, "tests", "_host_interp.so")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(REPO, "tests", "host_interp_harness.cpp")
    inc = os.path.join(REPO, "kmos_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in ("kb_interp.h", "kb_common.h")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", inc,
                               "-o", SO, src])
    L = C.CDLL(SO)
    i32p, i64p, f64p = (np.ctypeslib.ndpointer(dtype=d, flags="C_CONTIGUOUS") for d in (np.int32, np.int64, np.float64))
    L.kbh_create.restype = C.c_void_p
    L.kbh_create.argtypes = [i32p, C.c_int64, i32p, C.c_uint64, C.c_uint32]
    for f, a in [("destroy", []), ("set_rates", [f64p]), ("set_lut", [f64p]), ("init_state", [C.c_int]),
                 ("set_configuration", [i32p, C.c_int]), ("do_steps", [C.c_int64]), ("get_lattice", [i32p]),
                 ("get_procstat", [i64p]), ("get_nsites", [i32p]), ("get_integ", [f64p]), ("get_avail", [i32p]),
                 ("kmc_time", []), ("kmc_step", [])]:
        getattr(L, "kbh_" + f).argtypes = [C.c_void_p] + a
    L.kbh_kmc_time.restype = C.c_double
    L.kbh_kmc_step.restype = C.c_int64
    return L


CASES = [
    ("ab_local_smart", [20, 20], 5000),
    ("ab_lat_int", [10, 12], 3000),
    ("ab_otf", [10, 12], 3000),
    ("mini_101_local_smart", [20, 20], 2000),
    ("zgb_local_smart", [16, 16], 5000),
    ("zgb_lat_int", [9, 7], 3000),
    ("ruo2_local_smart", [20, 20], 5000),
    ("ruo2_lat_int", [6, 6], 2000),
    ("pairwise_lat_int", [12, 12], 3000),
    ("pairwise_local_smart", [8, 9], 2000),
    ("pairwise_otf_otf", [12, 10], 3000),
    ("pdopd_local_smart", [6, 5], 4000),  # multi-lattice: create_/annihilate_ routines, null_species = 4
    ("pdopd_lat_int", [6, 5], 4000),
    ("pairwise84_lat_int", [9, 8], 3000),
    ("ruo2default_otf", [8, 7], 3000),
    ("intzgb_otf", [10, 9], 3000),
    ("hop3d_local_smart", [5, 4, 3], 3000),  # our own 3-d / 1-d hop models: z and 1-d index arithmetic
    ("hop3d_lat_int", [4, 3, 5], 3000),
    ("hop3d_otf", [3, 5, 4], 3000),
    ("hop1d_local_smart", [17], 2000),
    ("hop1d_lat_int", [23], 2000),
    # further reference examples (round 2): H/Pt(111) with two hollow sites per cell, the reference's 1-d model
    # (Lotka-Volterra, diffusion and sand pile dead-lock from the default state: GPU parity runs them from
    # random configurations, tests/test_gpu_parity.py)
    ("pt111_local_smart", [8, 7], 3000), ("pt111_lat_int", [7, 6], 3000),
    ("einsd_local_smart", [23], 2000), ("einsd_lat_int", [19], 2000),
    # multidentate adsorbates (examples/multidentate.py): one species spans two and four sites
    ("multidentate_local_smart", [9, 8], 3000), ("multidentate_lat_int", [8, 7], 3000),
    ("multidentate_otf", [8, 7], 3000),
    ("zgb_otf", [12, 10], 3000), ("pt111_otf", [8, 7], 3000), ("einsd_otf", [23], 2000),   # 1-d otf
]


@pytest.mark.parametrize("name,size,steps", CASES)
def test_host_build_of_cuda_interpreter_matches_oracle(harness, name, size, steps):
    L = harness
    ir, blob, info = load_model(name, with_device=False)
    rng = np.random.RandomState(3)
    rates = np.exp(rng.uniform(-1.0, 1.0, len(ir["procs"])))
    lut = otf_mod.build_lut(ir, info, rates) if ir["backend"] == "otf" else None
    o = oracle.Oracle(blob, size, seed=42, replica=5, rates=rates, lut=lut)
    size3 = np.ones(3, dtype=np.int32)
    size3[:len(size)] = size
    h = L.kbh_create(blob, blob.size, size3, 42, 5)
    assert h
    L.kbh_set_rates(h, np.ascontiguousarray(rates))
    if lut is not None:
        L.kbh_set_lut(h, np.ascontiguousarray(lut))
    assert L.kbh_init_state(h, int(blob[9])) == 0

    def compare():
        lat = np.zeros(o.volume, dtype=np.int32); L.kbh_get_lattice(h, lat)
        assert np.array_equal(lat, o.lattice)
        ps = np.zeros(o.n_proc, dtype=np.int64); L.kbh_get_procstat(h, ps)
        assert np.array_equal(ps, o.procstat)
        ns = np.zeros(o.n_proc, dtype=np.int32); L.kbh_get_nsites(h, ns)
        assert np.array_equal(ns, o.nr_of_sites)
        av = np.zeros((o.n_proc, o.volume, 2), dtype=np.int32); L.kbh_get_avail(h, av)
        assert np.array_equal(av, o.avail_sites)
        assert L.kbh_kmc_step(h) == o.kmc_step
        assert abs(L.kbh_kmc_time(h) - o.kmc_time) <= 1e-12 * abs(o.kmc_time)
        ig = np.zeros(o.n_proc); L.kbh_get_integ(h, ig)
        np.testing.assert_allclose(ig, o.integ_rates, rtol=1e-12, atol=0)

    compare()
    for chunk in range(5):
        assert o.do_steps(steps // 5) == 0
        assert L.kbh_do_steps(h, steps // 5) == 0
        compare()
    # set_configuration + adjust_database on a random lattice
    spec = rng.randint(0, len(ir["species"]), o.volume).astype(np.int32)
    assert o.set_configuration(spec) == 0
    assert L.kbh_set_configuration(h, spec, int(blob[9])) == 0
    compare()
    L.kbh_destroy(h)


def _interval_search_by_the_book(arr, value):
    """base.interval_search_real transliterated statement by statement (kmos/fortran_src/base.mpy:1234-1338),
    1-based; returns 0 where the Fortran prints its dead-lock message and stops (or walks off the array)."""
    n = len(arr)
    left, right = 1, n
    while True:
        mid = (right + left) >> 1          # ISHFT(right+left, -1), computed before the exit test
        if left >= right:
            break
        if value < arr[mid - 1]:
            right = mid
        else:
            left = mid + 1
    if arr[mid - 1] == 0.0:                # nonzerosearch
        while True:
            if mid > n:
                return 0
            if arr[mid - 1] > 0.0:
                if mid >= n:
                    return 0
                break
            mid += 1
    while mid != 1 and arr[mid - 2] >= arr[mid - 1]:   # leftmostsearch
        mid -= 1
    return mid


def test_interval_search_real_edge_cases(harness):
    """The reference has no unit test for this routine although its docstring asks for one (base.mpy:1253-1255).
    Three implementations must agree: the transliteration above, the oracle's C, and the routine the CUDA kernels
    compile (KbInterp::interval_search_real, built for the host) -- on plateaus (zero-rate processes), leading
    and trailing zeros, values exactly on an entry, value == last entry (ran = 1 cannot occur, but 0 can)."""
    harness.kbh_interval_search_real.argtypes = [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_int, C.c_double]
    harness.kbh_interval_search_real.restype = C.c_int
    rng = np.random.RandomState(12)
    cases = [([1.0], [0.0, 0.5, 1.0]), ([0.0, 0.0, 2.0, 2.0, 5.0], [0.0, 1.9, 2.0, 4.999, 5.0]),
             ([3.0, 3.0, 3.0], [0.0, 2.9, 3.0]), ([0.0, 0.0, 0.0], [0.0]), ([0.0, 0.0, 7.0], [0.0, 6.0]),
             ([1.0, 1.0, 4.0, 4.0, 4.0, 9.0, 9.0], [0.0, 0.999, 1.0, 3.9, 4.0, 8.9, 9.0])]
    for _ in range(300):
        n = rng.randint(1, 40)
        inc = rng.choice([0.0, 0.0, 1.0, 0.25, 3.0], size=n) * rng.uniform(0.5, 2.0, size=n)
        arr = np.cumsum(inc)
        vals = list(rng.uniform(0, 1, 6) * arr[-1]) + list(arr[rng.randint(0, n, 3)]) + [0.0]
        cases.append((list(arr), vals))
    checked = 0
    for arr, vals in cases:
        a = np.asarray(arr, dtype=np.float64)
        for v in vals:
            want = _interval_search_by_the_book(arr, v)
            assert oracle.interval_search_real(a, v) == want, (arr, v)
            assert harness.kbh_interval_search_real(a, a.size, float(v)) == want, (arr, v)
            checked += 1
    assert checked > 2500
