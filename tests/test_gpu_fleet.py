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
