"""ModelRunner: grid construction as in the reference, and run(gpus=N) -- grid points dealt to N worker
processes (one per GPU), rows gathered in grid order, one .dat file -- against the unsharded scan.  Uses
tests/fake_model.py so that it runs without a GPU; tests/test_gpu_multi.py repeats it on two real GPUs."""
import numpy as np

from kmos_b200 import runner

import fake_model


class Scan(runner.ModelRunner):
    T = runner.TemperatureParameter(min=450, max=650, steps=5)
    p_COgas = runner.PressureParameter(min=1e-2, max=1e2, steps=3)


def test_grids_follow_the_reference():
    assert repr(runner.PressureParameter(1)) == "[pressure] min: 1, max: 1, steps: 1"
    np.testing.assert_allclose(runner.PressureParameter(min=1, max=100, steps=3).get_grid(), [1, 10, 100])
    g = runner.TemperatureParameter(min=400, max=800, steps=3).get_grid()        # regular in 1/T
    np.testing.assert_allclose(1.0 / g, np.linspace(1 / 400.0, 1 / 800.0, 3))
    np.testing.assert_allclose(runner.LogParameter(min=-1, max=1, steps=3).get_grid(), [0.1, 1, 10])
    pts = Scan("m.json").grid_points()
    assert len(pts) == 15 and list(pts[0]) == ["T", "p_COgas"] and pts[1]["T"] == pts[0]["T"]


def test_run_on_several_gpus_gathers_the_same_rows(tmp_path):
    kw = dict(init_steps=100, sample_steps=50, samples=1, random_seed=7)
    one = Scan("m.json", seeds=2, model_factory=fake_model.FakeModel, name="one")
    h1, rows1 = one.run(outfile=str(tmp_path / "one.dat"), per_replica=True, **kw)
    three = Scan("m.json", seeds=2, model_factory=fake_model.FakeModel, name="three")
    h3, rows3 = three.run(outfile=str(tmp_path / "three.dat"), per_replica=True, gpus=3, **kw)
    assert h1 == h3 and rows1.shape == (30, 4)
    assert np.array_equal(rows1, rows3)                      # global Philox keys: sharding changes nothing
    assert np.array_equal(rows1[:, 2], (7 + np.arange(30)) * 1e-3)
    a, b = open(tmp_path / "one.dat").read(), open(tmp_path / "three.dat").read()
    assert a == b and a.startswith("#T p_COgas tof kmc_steps\n# T = ") and len(a.splitlines()) == 1 + 2 + 2 + 30
    # one process, the GPUs driven through a fleet: the same rows and file again
    fleet = Scan("m.json", seeds=2, model_factory=fake_model.FakeModel, name="fleet")
    hf, rowsf = fleet.run(outfile=str(tmp_path / "fleet.dat"), per_replica=True, gpu_ids=[0, 1, 2], **kw)
    assert hf == h1 and np.array_equal(rowsf, rows1) and open(tmp_path / "fleet.dat").read() == a
    # means over the seeds of a grid point
    _h, rows = Scan("m.json", seeds=2, model_factory=fake_model.FakeModel).run(
        outfile=str(tmp_path / "mean.dat"), gpus=2, **kw)
    assert rows.shape == (15, 4) and np.allclose(rows[:, 2], (7 + 2 * np.arange(15) + 0.5) * 1e-3)
