"""kmos_b200_fleet_*: the replicas of one model dealt to several GPUs of one process (SURVEY 8b `gpu_ids[]`).

A device may be named more than once, so the sharding is exercised on a one-GPU box too; with two or more GPUs the
same cases also run across devices.  A fleet must give every replica the trajectory the oracle gives it -- hence the
trajectory a single batch gives it -- whatever the number of shards."""
import numpy as np
import pytest

from conftest import load_model
from kmos_b200 import capi
from util import compare_batch, make_inputs, run_oracles

pytestmark = pytest.mark.gpu


def _gpu_lists():
    n = capi.lib().kmos_b200_device_count()
    lists = [[0, 0, 0]]
    if n >= 2:
        lists.append([0, 1])
    return lists


@pytest.mark.parametrize("name,size,chunks", [
    ("ruo2_local_smart", [8, 8], [300, 700]),      # generated kernel, one module attached per shard
    ("pairwise_lat_int", [8, 8], [400]),           # warp-per-replica kernel
    ("pairwise_otf_otf", [8, 8], [300]),               # otf: per-replica gr_<proc> tables are sharded with the replicas
])
def test_fleet_matches_oracle_and_single_batch(name, size, chunks):
    from kmos_b200 import engine
    ir, blob, info = load_model(name)
    model = engine.Model(ir=ir, blob=blob, info=info)
    R = 10  # not divisible by three shards: 3 + 3 + 4
    rates, lut, seeds = make_inputs(ir, info, R, seed=3)
    single = engine.Batch(model, R, size, seeds=seeds, rates=rates, lut=lut)
    for n in chunks:
        single.do_steps(n)
    groups = np.arange(R, dtype=np.int32) % 2
    t_single = single.reduce_tallies(groups, 2)
    for gpu_ids in _gpu_lists():
        fleet = engine.Fleet(model, R, size, gpu_ids=gpu_ids, seeds=seeds, rates=rates, lut=lut)
        assert len(fleet.shards) == len(gpu_ids)
        assert fleet.bounds == [(R * k // len(gpu_ids), R * (k + 1) // len(gpu_ids)) for k in range(len(gpu_ids))]
        gen = run_oracles(blob, size, rates, lut, seeds, chunks)
        compare_batch(fleet, next(gen), avail_replicas=(0, 4, 9))
        for n in chunks:
            fleet.do_steps(n)
            compare_batch(fleet, next(gen), avail_replicas=(0, 4, 9))
        # bit for bit the single batch's replicas
        assert np.array_equal(fleet.kmc_time, single.kmc_time)
        assert np.array_equal(fleet.integ_rates, single.integ_rates)
        assert np.array_equal(fleet.procstat, single.procstat) and np.array_equal(fleet.lattice, single.lattice)
        # tallies: counts exact, f64 sums to rounding (the partial sums are added in a different order)
        t = fleet.reduce_tallies(groups, 2)
        a, b = fleet.split_tally(t), single.split_tally(t_single)
        for key in ("procstat", "kmc_steps", "n_replicas", "occupation"):
            assert np.array_equal(a[key], b[key]), key
        np.testing.assert_allclose(a["kmc_time"], b["kmc_time"], rtol=1e-13)
        np.testing.assert_allclose(a["integ_rates"], b["integ_rates"], rtol=1e-13)
        fleet.close()
    single.close()


def test_fleet_fewer_replicas_than_gpus_and_errors():
    from kmos_b200 import engine
    ir, blob, info = load_model("mini_101_local_smart")
    model = engine.Model(ir=ir, blob=blob, info=info)
    rates, lut, seeds = make_inputs(ir, info, 2, seed=1)
    fleet = engine.Fleet(model, 2, [6, 6], gpu_ids=[0, 0, 0, 0, 0], seeds=seeds, rates=rates)
    assert len(fleet.shards) == 2 and [hi - lo for lo, hi in fleet.bounds] == [1, 1]   # three GPUs stay idle
    fleet.do_steps(500)
    gen = run_oracles(blob, [6, 6], rates, None, seeds, [500])
    next(gen)
    compare_batch(fleet, next(gen), avail_replicas=(0, 1))
    assert list(fleet.split_tally(fleet.reduce_tallies())["kmc_steps"]) == [1000]
    fleet.close()
    with pytest.raises(capi.KmosB200Error, match="device 99"):
        engine.Fleet(model, 4, [6, 6], gpu_ids=[0, 99])
    with pytest.raises(capi.KmosB200Error):
        engine.Fleet(model, 4, [6, 6], gpu_ids=[])


def test_kmc_model_on_a_fleet_gives_the_single_gpu_rows(tmp_path):
    """KMC_Model(gpu_ids=...) -- the front-end of one process on several GPUs: the reference-shaped outputs, the
    replay entry points and the configuration round trip are those of the same replicas on one batch (which
    tests/test_gpu_model.py holds against the oracle)."""
    import os
    from conftest import GOLDEN
    from kmos_b200.model import KMC_Model
    path = os.path.join(GOLDEN, "models", "ab_local_smart.json")
    points = [{"p_COgas": 1.0, "p_O2gas": 1.0}, {"p_COgas": 0.3, "p_O2gas": 2.0}, {"p_COgas": 3.0, "p_O2gas": 0.5},
              {"p_COgas": 2.0, "p_O2gas": 0.7}, {"p_COgas": 0.5, "p_O2gas": 0.5}]
    n_dev = capi.lib().kmos_b200_device_count()
    kw = dict(size=[12, 10], n_replicas=5, parameters=points, random_seed=5)
    with KMC_Model(path, **kw) as one, KMC_Model(path, gpu_ids=[0, 1 % n_dev, 0], **kw) as many:
        assert [hi - lo for lo, hi in many.batch.bounds] == [1, 2, 2]
        for m in (one, many):
            m.do_steps(2000)
        a = one.get_std_sampled_data_all(samples=3, sample_size=3000, tof_method="integ")
        b = many.get_std_sampled_data_all(samples=3, sample_size=3000, tof_method="integ")
        assert np.array_equal(np.asarray(a), np.asarray(b))
        # get_next_kmc_step / run_proc_nr (the replay loop of tests/test_run/test_run.py) across shards
        for _ in range(20):
            pa, sa = one.batch.get_next_kmc_step()
            pb, sb = many.batch.get_next_kmc_step()
            assert np.array_equal(pa, pb) and np.array_equal(sa, sb)
            one.batch.run_proc_nr(pa, sa)
            many.batch.run_proc_nr(pb, sb)
        assert np.array_equal(one.batch.lattice, many.batch.lattice)
        # configuration of a replica in the last shard into one of the first; set_rate_const on one replica
        many.dump_config(str(tmp_path / "cfg"), replica=4)
        many.load_config(str(tmp_path / "cfg"), replica=0)
        one.dump_config(str(tmp_path / "cfg1"), replica=4)
        one.load_config(str(tmp_path / "cfg1"), replica=0)
        for m in (one, many):
            m.batch.set_rate_const(1, 0.125, replica=3)
            m.batch.set_kmc_time(np.arange(5.0))
            m.do_steps(1500)
        assert many.batch.rates[3, 0] == 0.125 and many.batch.rates[2, 0] != 0.125
        for name in ("lattice", "procstat", "kmc_time", "kmc_step", "nr_of_sites", "accum_rates", "kmc_time_step"):
            assert np.array_equal(getattr(one.batch, name), getattr(many.batch, name)), name
        assert np.array_equal(one.batch.avail_sites(4), many.batch.avail_sites(4))
        # restart file of a replica that lives on the second shard
        many.batch.save_system(str(tmp_path / "r.reload"), replica=2)
        many.do_steps(700)
        ref = many.batch.lattice[2].copy()
        many.batch.reload_system(str(tmp_path / "r.reload"), replica=2)
        many.do_steps(700)
        assert np.array_equal(many.batch.lattice[2], ref)


def test_model_runner_on_a_fleet_writes_the_same_dat_file(tmp_path):
    """ModelRunner.run(gpu_ids=[...]): the scan as one fleet of this process -- byte for byte the .dat file of the
    single-GPU scan (kmos/run/__init__.py:2129-2139 format)."""
    import os
    from conftest import GOLDEN
    from kmos_b200 import runner

    class Scan(runner.ModelRunner):
        T = runner.TemperatureParameter(min=500, max=600, steps=3)
        p_COgas = runner.PressureParameter(min=0.5, max=5, steps=3)

    model = os.path.join(GOLDEN, "models", "ruo2_local_smart.json")
    n_dev = capi.lib().kmos_b200_device_count()
    kw = dict(init_steps=3000, sample_steps=3000, samples=2, random_seed=11)
    h1, r1 = Scan(model, size=8, seeds=3, name="a").run(outfile=str(tmp_path / "a.dat"), **kw)
    h2, r2 = Scan(model, size=8, seeds=3, name="b").run(outfile=str(tmp_path / "b.dat"), gpu_ids=[0, 1 % n_dev, 0, 0], **kw)
    assert h1 == h2 and np.array_equal(r1, r2)
    assert open(tmp_path / "a.dat").read() == open(tmp_path / "b.dat").read()
