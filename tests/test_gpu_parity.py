"""GPU parity: the CUDA engines, called through the C-ABI, against the CPU oracle on the same Philox stream.

Bit-exact: lattice, procstat, nr_of_sites, both planes of avail_sites, kmc_step, status.
Floating point: kmc_time within 1e-12 relative (north_star), integ_rates within 1e-10 relative
(device log() vs glibc log() may differ in the last ulp of each time increment).
"""
import numpy as np
import pytest

from conftest import load_model
from kmos_b200 import capi
from util import compare_batch, make_inputs, run_oracles

pytestmark = pytest.mark.gpu


def _engine():
    from kmos_b200 import engine
    return engine


SMEM_CASES = [
    ("mini_101_local_smart", [20, 20], 16, [1, 999, 3000]),
    ("ab_local_smart", [20, 20], 24, [500, 2500, 3000]),
    ("zgb_local_smart", [16, 12], 13, [1000, 4000]),
    ("zgb_local_smart", [64, 64], 3, [3000, 3000]),
    ("ruo2_local_smart", [20, 20], 32, [2000, 4000, 6000]),
    ("ruo2_local_smart", [5, 7], 7, [3000, 3000]),
    ("pairwise_local_smart", [10, 9], 9, [2000, 2000]),
    ("hop3d_local_smart", [5, 6, 5], 7, [2000, 2000]),   # 3-d and 1-d lattices (no reference example has them)
    ("hop3d_local_smart", [9, 8, 7], 4, [3000]),
    ("hop1d_local_smart", [17], 6, [2000, 2000]),
    # further examples of the reference (examples/render_{Lotka_Volterra_model,diffusion_model,sand_model,Pt_111,
    # einsD}.py)
    ("pt111_local_smart", [8, 7], 9, [3000, 3000]),
    ("einsd_local_smart", [23], 9, [2000, 2000]),
    ("multidentate_local_smart", [9, 8], 9, [3000, 3000]),   # examples/multidentate.py: 2- and 4-site species
]


KINDS = {"smem": capi.KERNEL_SMEM, "generic": capi.KERNEL_GENERIC, "warp_hbm": capi.KERNEL_WARP_HBM,
         "generated": capi.KERNEL_GENERATED}


@pytest.mark.parametrize("name,size,R,chunks", SMEM_CASES)
@pytest.mark.parametrize("kernel", ["generated", "smem", "generic", "warp_hbm"])
def test_local_smart_parity(name, size, R, chunks, kernel):
    """Every local_smart fixture on every kernel that takes it; "generated" = the exporter-emitted per-model
    module (kmos_b200/codegen.py), lane-group width chosen by the generator."""
    engine = _engine()
    ir, blob, info = load_model(name)
    rates, lut, seeds = make_inputs(ir, info, R, seed=len(name))
    model = engine.Model(ir=ir, blob=blob, info=info)
    kind = KINDS[kernel]
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates, kernel=kind)
    assert batch.kernel_info()["kernel_name"] == kernel
    gen = run_oracles(blob, size, rates, lut, seeds, chunks)
    compare_batch(batch, next(gen), avail_replicas=range(min(R, 3)))
    for n, oracles in zip(chunks, gen):
        batch.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(0, R - 1))
    batch.close()


@pytest.mark.parametrize("lpr", [4, 8, 16, 32])
@pytest.mark.parametrize("name,size,R,chunks", [
    ("ruo2_local_smart", [20, 20], 7, [3000, 3000]),   # 7 replicas: the last team of 2 / 4 is incomplete
    ("ruo2_local_smart", [5, 7], 9, [3000, 3000]),
    ("zgb_local_smart", [16, 12], 13, [1000, 4000]),
    ("pairwise_local_smart", [10, 9], 5, [2000, 2000]),
    ("ab_local_smart", [20, 20], 33, [500, 2500]),
    ("hop3d_local_smart", [5, 6, 5], 6, [2000]),
    ("mini_101_local_smart", [6, 5], 19, [1000, 1000]),  # 19 replicas: three teams of eight at 4 lanes, the last incomplete
    ("multidentate_local_smart", [9, 8], 9, [2000]),
])
def test_generated_kernel_lane_group_widths(name, size, R, chunks, lpr):
    """The generated kernel steps 32/lpr replicas per warp in lock step; every width must walk the oracle's
    trajectory, with replica counts that leave the last team incomplete."""
    engine = _engine()
    from kmos_b200 import codegen
    ir, blob, info = load_model(name)
    if lpr not in codegen.lane_group_widths(len(ir["procs"])):
        pytest.skip("groups of %d lanes are for models with at most 16 processes" % lpr)
    rates, lut, seeds = make_inputs(ir, info, R, seed=len(name) + lpr)
    model = engine.Model(ir=ir, blob=blob, info=info)
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates, kernel=capi.KERNEL_GENERATED, lpr=lpr)
    assert batch.kernel_info()["kernel_name"] == "generated"
    gen = run_oracles(blob, size, rates, lut, seeds, chunks)
    compare_batch(batch, next(gen), avail_replicas=(0,))
    for n, oracles in zip(chunks, gen):
        batch.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(0, R - 1))
    batch.close()


@pytest.mark.parametrize("kernel", ["generic", "warp_hbm"])
@pytest.mark.parametrize("name,size", [("pdopd_local_smart", [6, 5]), ("pairwise84_local_smart", [9, 8])])
def test_local_smart_models_beyond_the_lane_tables(name, size, kernel):
    """Models the shared-memory kernel declines: Pd(100)/PdO of the reference's export tests (two lattices, 25
    sites per cell, create_/annihilate_ routines, a declared null_species) and an 84-process pairwise model
    (more than 64 processes: 4 process segments per lane).  The two HBM-state kernels must match the oracle."""
    engine = _engine()
    ir, blob, info = load_model(name)
    R, chunks = 6, [3000, 3000]
    rates, lut, seeds = make_inputs(ir, info, R, seed=11)
    model = engine.Model(ir=ir, blob=blob, info=info)
    auto = engine.Batch(model, R, size, seeds=seeds, rates=rates)
    assert auto.kernel_info()["kernel_name"] == "warp_hbm"
    auto.close()
    kind = capi.KERNEL_WARP_HBM if kernel == "warp_hbm" else capi.KERNEL_GENERIC
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates, kernel=kind)
    gen = run_oracles(blob, size, rates, lut, seeds, chunks)
    compare_batch(batch, next(gen), avail_replicas=(0,))
    for n, oracles in zip(chunks, gen):
        batch.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(0, R - 1))
    batch.close()


LATINT_CASES = [
    ("hop3d_lat_int", [5, 7, 6], 5, [2000, 2000]),
    ("hop1d_lat_int", [23], 5, [2000, 2000]),
    ("pdopd_lat_int", [6, 5], 5, [3000, 3000]),
    ("pairwise84_lat_int", [12, 11], 6, [3000, 3000]),
    ("ab_lat_int", [10, 12], 6, [1500, 1500]),
    ("mini_101_lat_int", [7, 5], 5, [700, 700]),
    ("zgb_lat_int", [12, 12], 5, [2000, 2000]),
    ("ruo2_lat_int", [8, 8], 5, [1500, 1500]),
    ("pairwise_lat_int", [16, 16], 8, [2000, 2000]),
    ("pairwise_lat_int", [5, 3], 3, [1000]),
    ("pt111_lat_int", [7, 6], 5, [2000, 2000]),
    ("einsd_lat_int", [19], 5, [2000, 2000]),
    ("multidentate_lat_int", [8, 7], 5, [2000, 2000]),
]


@pytest.mark.parametrize("name,size,R,chunks", LATINT_CASES)
@pytest.mark.parametrize("kernel", ["warp_hbm", "generic"])
def test_lat_int_parity(name, size, R, chunks, kernel):
    engine = _engine()
    ir, blob, info = load_model(name)
    rates, lut, seeds = make_inputs(ir, info, R, seed=len(name) + 1)
    model = engine.Model(ir=ir, blob=blob, info=info)
    kind = capi.KERNEL_WARP_HBM if kernel == "warp_hbm" else capi.KERNEL_GENERIC
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates, kernel=kind)
    assert batch.kernel_info()["kernel_name"] == kernel
    gen = run_oracles(blob, size, rates, lut, seeds, chunks)
    compare_batch(batch, next(gen), avail_replicas=(0,))
    for n, oracles in zip(chunks, gen):
        batch.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(0, R - 1))
    batch.close()


@pytest.mark.parametrize("name,size,R,chunks", LATINT_CASES + [("pairwise_lat_int", [24, 20], 6, [3000, 3000])])
def test_lat_int_interior_cells(name, size, R, chunks, monkeypatch):
    """The warp-HBM kernel addresses the probes of events away from the lattice edges without the periodic
    arithmetic (cell + linear offset).  The planner enables that for mostly-interior lattices only (128x128);
    forced on here so that small lattices mix interior and edge events, 1-d to 3-d."""
    engine = _engine()
    monkeypatch.setenv("KMOS_B200_INTERIOR", "1")
    ir, blob, info = load_model(name)
    rates, lut, seeds = make_inputs(ir, info, R, seed=len(name) + 7)
    model = engine.Model(ir=ir, blob=blob, info=info)
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates, kernel=capi.KERNEL_WARP_HBM)
    assert batch.kernel_info()["kernel_name"] == "warp_hbm"
    gen = run_oracles(blob, size, rates, lut, seeds, chunks)
    next(gen)
    for n, oracles in zip(chunks, gen):
        batch.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(0, R - 1))
    batch.close()


OTF_CASES = [
    ("ab_otf", [10, 12], 6, [1500, 1500]),
    ("pairwise_otf_otf", [16, 16], 8, [2000, 2000]),
    ("mini_101_otf", [6, 6], 4, [500, 500]),
    ("hop3d_otf", [5, 6, 5], 5, [1500, 1500]),
    ("ruo2default_otf", [8, 7], 5, [1500, 1500]),   # the reference's committed otf export: 36 processes, 2 sites/cell
    ("intzgb_otf", [10, 9], 6, [2000, 2000]),        # interacting ZGB: bystander-dependent rates (1150 LUT entries)
    ("multidentate_otf", [8, 7], 5, [2000, 2000]),
    ("zgb_otf", [12, 10], 5, [2000, 2000]),
    ("pt111_otf", [8, 7], 5, [2000, 2000]),
    ("einsd_otf", [23], 5, [2000, 2000]),            # a 1-d otf model
]


@pytest.mark.parametrize("name,size,R,chunks", OTF_CASES + [("pairwise_otf_otf", [40, 36], 5, [1500, 1500])])
@pytest.mark.parametrize("kernel", ["warp_hbm", "generic"])
def test_otf_parity(name, size, R, chunks, kernel):
    engine = _engine()
    ir, blob, info = load_model(name)
    rates, lut, seeds = make_inputs(ir, info, R, seed=len(name) + 1)
    model = engine.Model(ir=ir, blob=blob, info=info)
    kind = capi.KERNEL_WARP_HBM if kernel == "warp_hbm" else capi.KERNEL_GENERIC
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates, lut=lut, kernel=kind)
    assert batch.kernel_info()["kernel_name"] == kernel
    gen = run_oracles(blob, size, rates, lut, seeds, chunks)
    compare_batch(batch, next(gen), avail_replicas=(0,))
    for n, oracles in zip(chunks, gen):
        batch.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(0, R - 1))
    batch.close()


def test_local_smart_lattice_too_large_for_shared_memory():
    """128x128 cells exceed the 13-bit positions of the compact class entries: the planner falls back to the
    HBM-resident warp kernel, which must still be bit-exact."""
    engine = _engine()
    ir, blob, info = load_model("zgb_local_smart")
    R, size, n = 6, [128, 128], 3000
    rates, lut, seeds = make_inputs(ir, info, R, seed=77)
    model = engine.Model(ir=ir, blob=blob, info=info)
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates)
    assert batch.kernel_info()["kernel_name"] == "warp_hbm"
    gen = run_oracles(blob, size, rates, lut, seeds, [n])
    next(gen)
    batch.do_steps(n)
    compare_batch(batch, next(gen), avail_replicas=(0, R - 1))
    batch.close()


@pytest.mark.parametrize("kernel", ["generated", "smem", "warp_hbm", "generic"])
@pytest.mark.parametrize("name,size", [
    ("lotka_local_smart", [12, 10]), ("diffusion_local_smart", [9, 9]), ("sand_local_smart", [10, 8]),
    ("lotka_lat_int", [9, 8]), ("diffusion_lat_int", [8, 7]), ("sand_lat_int", [8, 8]),
])
def test_reference_examples_from_random_configurations(name, size, kernel):
    """examples/render_{Lotka_Volterra_model,diffusion_model,sand_model}.py: nothing can happen on their default
    lattice (the examples' users set a configuration first), so every replica starts from its own random
    configuration (set_configuration + _adjust_database, kmos/run/__init__.py:1411-1457).  Replicas whose
    populations die out must stop with the oracle's dead-lock status at the oracle's step."""
    engine = _engine()
    ir, blob, info = load_model(name)
    if ir["backend"] == "lat_int" and kernel in ("generated", "smem"):
        pytest.skip("local_smart kernels")
    R = 7
    rates, lut, seeds = make_inputs(ir, info, R, seed=len(name) + 2)
    batch = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, size, seeds=seeds, rates=rates, kernel=KINDS[kernel])
    assert batch.kernel_info()["kernel_name"] == kernel
    oracles = next(run_oracles(blob, size, rates, lut, seeds, []))
    rng = np.random.RandomState(7)
    spec = rng.randint(0, len(ir["species"]), (R, batch.volume)).astype(np.int32)
    batch.set_configuration(spec)
    for r, o in enumerate(oracles):
        assert o.set_configuration(spec[r]) == 0
    compare_batch(batch, oracles, avail_replicas=(0, R - 1))
    for n in (700, 1800):
        batch.do_steps(n)
        for o in oracles:
            o.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(0, R - 1))
    batch.close()


def test_set_configuration_then_steps():
    engine = _engine()
    ir, blob, info = load_model("ruo2_local_smart")
    R, size = 5, [8, 6]
    rates, lut, seeds = make_inputs(ir, info, R, seed=5)
    model = engine.Model(ir=ir, blob=blob, info=info)
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates)
    oracles = next(run_oracles(blob, size, rates, lut, seeds, []))
    rng = np.random.RandomState(1)
    spec = rng.randint(0, len(ir["species"]), (R, batch.volume)).astype(np.int32)
    batch.set_configuration(spec)
    for r, o in enumerate(oracles):
        assert o.set_configuration(spec[r]) == 0
    compare_batch(batch, oracles, avail_replicas=range(R))
    batch.do_steps(2500)
    for o in oracles:
        o.do_steps(2500)
    compare_batch(batch, oracles, avail_replicas=range(R))
    # one replica only
    spec1 = rng.randint(0, len(ir["species"]), batch.volume).astype(np.int32)
    batch.set_configuration(spec1, replica=3)
    assert oracles[3].set_configuration(spec1) == 0
    compare_batch(batch, oracles, avail_replicas=(2, 3))


def test_deadlock_is_a_status_not_a_hang():
    """All rate constants zero: the reference prints its dead-lock message and stops (base.mpy:1296-1305)."""
    engine = _engine()
    ir, blob, info = load_model("mini_101_local_smart")
    model = engine.Model(ir=ir, blob=blob, info=info)
    for kind in (capi.KERNEL_GENERATED, capi.KERNEL_SMEM, capi.KERNEL_GENERIC):
        batch = engine.Batch(model, 4, [6, 6], rates=np.zeros((4, 2)), kernel=kind)
        batch.do_steps(100)
        assert np.all(batch.status == capi.REPLICA_DEADLOCK)
        assert np.all(batch.kmc_step == 0)
        batch.close()
    # adsorption only: the lattice fills up after exactly 36 events, then nothing is available
    rates = np.array([[1.0, 0.0]] * 3)
    batch = engine.Batch(model, 3, [6, 6], rates=rates)
    batch.do_steps(100)
    assert np.all(batch.status == capi.REPLICA_DEADLOCK)
    assert np.all(batch.kmc_step == 36)
    assert np.all(batch.lattice == 0)  # CO everywhere (species sorted by name: CO=0, empty=1)


def test_species_mismatch_reports_the_reference_error_tuple():
    """replace_species finds another species than the rule expects: the reference prints (old, new, found, site,
    step) and stops (base.mpy:1205-1228; KMC_Model.post_mortem reads the tuple).  A deliberately inconsistent
    rule set (take_CO expects 'empty' where its condition guarantees CO) must stop every replica at its first
    desorption with the oracle's tuple, on every kernel."""
    import copy
    from kmos_b200 import tables
    from oracle import oracle
    engine = _engine()
    ir, _blob, _info = load_model("mini_101_local_smart")
    ir = copy.deepcopy(ir)
    ir["routines"]["take_CO_simple_cubic_hollow"][0] = ["replace", [0, 0, 0, 0], 1, 1]
    blob, info = tables.build_blob(ir)
    model = engine.Model(ir=ir, blob=blob, info=info)
    R, size = 6, [6, 5]
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(3)
    rates = np.tile(np.array([100.0, 100.0]), (R, 1))
    for kind in (capi.KERNEL_GENERATED, capi.KERNEL_SMEM, capi.KERNEL_WARP_HBM, capi.KERNEL_GENERIC):
        batch = engine.Batch(model, R, size, seeds=seeds, rates=rates, kernel=kind)
        batch.do_steps(200)
        status, err, step, lat = batch.status, batch.error_info, batch.kmc_step, batch.lattice
        for r in range(R):
            o = oracle.Oracle(blob, size, seed=int(seeds[r]), replica=r, rates=rates[r])
            assert o.do_steps(200) == capi.REPLICA_SPECIES_MISMATCH
            ost, oerr = o.status
            assert status[r] == ost == capi.REPLICA_SPECIES_MISMATCH, (kind, r)
            assert list(err[r]) == list(oerr), (kind, r, err[r], oerr)
            assert step[r] == o.kmc_step
            assert np.array_equal(lat[r], o.lattice)
        batch.close()


def test_philox_hook_matches_oracle():
    from oracle import oracle
    L = capi.lib()
    for seed, rep, step in [(1, 0, 0), (2**40 + 17, 123, 2**33 + 5), (99, 4000000000, 77)]:
        ref = oracle.philox_step(seed, rep, step)
        got = [L.kmos_b200_philox_next(seed, rep, step, s) for s in range(3)]
        assert list(ref) == got
        assert 0 < got[0] <= 1 and 0 <= got[1] < 1 and 0 <= got[2] < 1


def test_full_size_ruo2_batch_properties():
    """BASELINE headline shape (RuO2 20x20, 16384 replicas): size-independent invariants on every replica
    plus full oracle parity on a sample of replicas."""
    engine = _engine()
    ir, blob, info = load_model("ruo2_local_smart")
    R, size, n = 16384, [20, 20], 2000
    rates, lut, seeds = make_inputs(ir, info, R, seed=9)
    model = engine.Model(ir=ir, blob=blob, info=info)
    batch = engine.Batch(model, R, size, seeds=seeds, rates=rates)
    assert batch.kernel_info()["kernel_name"] == "generated"
    batch.do_steps(n)
    assert np.all(batch.status == 0)
    assert np.all(batch.kmc_step == n)
    ps = batch.procstat
    assert np.all(ps.sum(axis=1) == n)            # every step executed exactly one process
    occ = batch.occupation
    np.testing.assert_allclose(occ.sum(axis=1), 1.0, atol=1e-12)  # every site holds exactly one species
    assert np.all(np.diff(np.sort(batch.kmc_time)) >= 0) and np.all(batch.kmc_time > 0)
    sample = [0, 1, 4097, 16383]
    from oracle import oracle
    lat = batch.lattice
    ns = batch.nr_of_sites
    t = batch.kmc_time
    for r in sample:
        o = oracle.Oracle(blob, size, seed=int(seeds[r]), replica=r, rates=rates[r])
        o.do_steps(n)
        assert np.array_equal(lat[r], o.lattice)
        assert np.array_equal(ps[r], o.procstat)
        assert np.array_equal(ns[r], o.nr_of_sites)
        assert np.array_equal(batch.avail_sites(r), o.avail_sites)
        assert abs(t[r] - o.kmc_time) <= 1e-12 * o.kmc_time
    # tallies: one group per 1024 replicas, totals must equal the per-replica sums
    groups = np.arange(R) // 1024
    tall = batch.split_tally(batch.reduce_tallies(groups, 16))
    assert np.array_equal(tall["procstat"].sum(axis=0), ps.sum(axis=0).astype(np.float64))
    assert np.all(tall["n_replicas"] == 1024) and np.all(tall["kmc_steps"] == 1024 * n)


@pytest.mark.parametrize("handover", ["generated", "smem"])
def test_reference_golden_trajectory_replayed_on_the_gpu(handover):
    """The reference's own known-answer trajectory (tests/test_run/ref_procs_sites_local_smart.log: 10 000
    (proc, site) events of the AB model, 20x20) is executed on the GPU through the reference's replay interface
    (model.run_proc_nr(proc, site), tests/test_run/test_run.py:51-53): every event must be enabled on the GPU
    when it is due, and lattice, nr_of_sites and both planes of avail_sites must follow the oracle replaying the
    same log.  Afterwards the batch continues on the generated / the shared-memory kernel (canonical -> compact
    repack)."""
    import os
    from conftest import GOLDEN
    from kmos_b200 import rates as rates_mod
    from oracle import oracle
    engine = _engine()
    ref = np.load(os.path.join(GOLDEN, "ab_ref_procs_sites.npy"))
    ir, blob, info = load_model("ab_local_smart")
    r = np.asarray(rates_mod.model_rates(ir))
    R = 2
    batch = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, [20, 20], seeds=np.array([1, 2], np.uint64),
                         rates=np.tile(r, (R, 1)))
    o = oracle.Oracle(blob, [20, 20], seed=1, replica=0, rates=r)
    V, P = batch.volume, len(ir["procs"])
    for i, (proc, site) in enumerate(ref):
        if i % 1000 == 0 or i == len(ref) - 1:
            av = batch.avail_sites(1)
            assert np.array_equal(av, o.avail_sites), "avail_sites differ before event %d" % i
            assert av.reshape(P, V, 2)[proc - 1, site - 1, 1] != 0, "event %d is not enabled on the GPU" % i
        batch.run_proc_nr(int(proc), int(site))
        o.run_proc_nr(int(proc), int(site))
    assert np.all(batch.status == 0) and np.all(batch.kmc_step == 0)
    for rep in range(R):
        assert np.array_equal(batch.lattice[rep], o.lattice)
        assert np.array_equal(batch.procstat[rep], o.procstat)
        assert np.array_equal(batch.nr_of_sites[rep], o.nr_of_sites)
        assert np.array_equal(batch.avail_sites(rep), o.avail_sites)
    assert batch.procstat[0].sum() == len(ref)
    # get_next_kmc_step agrees with the oracle's (same Philox draw, site selected with ran_time), then both step
    gp, gs = batch.get_next_kmc_step()
    op_, os_, st = o.get_next_kmc_step()
    assert st == 0 and (int(gp[0]), int(gs[0])) == (op_, os_)
    batch.select_kernel(KINDS[handover])
    assert batch.kernel_info()["kernel_name"] == handover
    batch.do_steps(2000)
    o.do_steps(2000)
    assert np.array_equal(batch.lattice[0], o.lattice)
    assert np.array_equal(batch.avail_sites(0), o.avail_sites)
    np.testing.assert_allclose(batch.kmc_time[0], o.kmc_time, rtol=1e-12)


@pytest.mark.parametrize("name,size,kernel", [
    ("ruo2_local_smart", [5, 7], "generated"), ("ab_local_smart", [4, 3], "generated"), ("zgb_local_smart", [6, 5], "generated"),
    ("ruo2_local_smart", [5, 7], "smem"), ("ruo2_local_smart", [5, 7], "warp_hbm"), ("ruo2_local_smart", [5, 7], "generic"),
    ("zgb_lat_int", [9, 7], "warp_hbm"), ("pairwise_otf_otf", [9, 8], "warp_hbm"), ("ab_local_smart", [4, 3], "smem"),
])
def test_many_small_launches_match_one_trajectory(name, size, kernel):
    """do_kmc_steps(1), (1), (2), ... must walk the same trajectory as the oracle: the persistent schedulers cut
    a launch into epochs and stage state in and out, so tiny and odd launch sizes are their edge cases (1 step,
    fewer steps than an epoch, fewer replicas than a CTA holds)."""
    engine = _engine()
    ir, blob, info = load_model(name)
    R = 3
    rates, lut, seeds = make_inputs(ir, info, R, seed=5)
    kind = KINDS[kernel]
    batch = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, size, seeds=seeds, rates=rates, lut=lut, kernel=kind)
    assert batch.kernel_info()["kernel_name"] == kernel
    chunks = [1, 1, 2, 3, 5, 17, 300, 1, 255, 256, 257]
    gen = run_oracles(blob, size, rates, lut, seeds, chunks)
    next(gen)
    for n, oracles in zip(chunks, gen):
        batch.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(R - 1,))
    batch.close()


@pytest.mark.parametrize("kernel", ["generated", "smem", "warp_hbm"])
def test_epochs_with_replicas_that_stop_midway(kernel, monkeypatch):
    """A launch cut into several epochs (forced here) while some replicas dead-lock after 36 events: their
    remaining work items must pass through without touching the state, the others keep stepping."""
    engine = _engine()
    monkeypatch.setenv("KMOS_B200_EPOCHS", "5")
    ir, blob, info = load_model("mini_101_local_smart")
    R, size = 40, [6, 6]
    rates = np.tile(np.array([3.0, 2.0]), (R, 1))
    rates[::3, 1] = 0.0  # adsorption only: the 36 sites fill up, then nothing is available
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(900)
    batch = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, size, seeds=seeds, rates=rates, kernel=KINDS[kernel])
    gen = run_oracles(blob, size, rates, None, seeds, [500, 123])
    next(gen)
    for n, oracles in zip([500, 123], gen):
        batch.do_steps(n)
        compare_batch(batch, oracles, avail_replicas=(0, 1))
    assert np.all(batch.status[::3] == capi.REPLICA_DEADLOCK) and np.all(batch.kmc_step[::3] == 36)
    assert np.all(batch.status[1::3] == 0) and np.all(batch.kmc_step[1::3] == 623)
    batch.close()
