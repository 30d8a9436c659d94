"""Multi-GPU correctness on hardware: replicas sharded over 2 NCCL ranks, tallies all-reduced
(int64 counts, float64 sums) must equal the tallies of the same replicas stepped on one GPU -- counts exactly,
sums to 1e-12.  Skipped with fewer than 2 GPUs (`gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_sharded_tallies_equal_single_gpu_tallies():
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "multi_gpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTI_GPU_OK world=2" in p.stdout


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_model_runner_on_two_gpus_writes_the_same_dat_file(tmp_path):
    """ModelRunner.run(gpus=2): grid points dealt to two GPUs, rows gathered on the parent -- identical to the
    single-GPU scan, byte for byte in the .dat file (kmos/run/__init__.py:2129-2139 format)."""
    import numpy as np
    from conftest import GOLDEN
    from kmos_b200 import runner

    class Scan(runner.ModelRunner):
        T = runner.TemperatureParameter(min=500, max=600, steps=3)
        p_COgas = runner.PressureParameter(min=0.5, max=5, steps=3)

    model = os.path.join(GOLDEN, "models", "ruo2_local_smart.json")
    kw = dict(init_steps=3000, sample_steps=3000, samples=2, random_seed=11)
    h1, r1 = Scan(model, size=8, seeds=3, name="a").run(outfile=str(tmp_path / "a.dat"), **kw)
    h2, r2 = Scan(model, size=8, seeds=3, name="b").run(outfile=str(tmp_path / "b.dat"), gpus=2, **kw)
    assert h1 == h2 and np.array_equal(r1, r2)
    assert open(tmp_path / "a.dat").read() == open(tmp_path / "b.dat").read()
