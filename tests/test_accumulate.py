"""update_accum_rate on the device (kb_smem.cuh / kb_latint.cuh): every lane adds the *packed non-zero* products of
its segment behind a run of zeros, in tiers of 8, 16 or 32 additions.  This must reproduce the reference's serial
float64 recurrence accum(i) = accum(i-1) + n(i)*r(i) (kmos/fortran_src/base.mpy:603-623) bit for bit; the argument
(x + 0.0 == x exactly) is checked here on the host with the same index arithmetic the kernels use.
"""
import numpy as np


def serial_prefix(x):
    acc, out = np.float64(0.0), []
    for v in x:
        acc = acc + np.float64(v)
        out.append(acc)
    return np.array(out)


def packed_tier_prefix(x):
    """One 32-entry segment as the warp computes it: Z = [32 zeros][packed non-zero products]."""
    assert len(x) == 32
    z = np.zeros(64)
    nz = [i for i in range(32) if x[i] != 0.0]
    for rank, i in enumerate(nz):
        z[32 + rank] = x[i]  # lane i writes at popc(nz & lanes < i)
    c = len(nz)
    tier = 8 if c <= 8 else 16 if c <= 16 else 32
    out = []
    for lane in range(32):
        k = sum(1 for i in nz if i <= lane)  # popc(nz & lanes <= lane)
        top = 32 + k
        acc = np.float64(0.0)
        for t in range(top - tier, top):
            acc = acc + z[t]
        out.append(acc)
    return np.array(out)


def test_packed_chain_equals_serial_recurrence():
    rng = np.random.RandomState(5)
    for trial in range(400):
        n_nonzero = rng.randint(0, 33)
        x = np.zeros(32)
        idx = rng.choice(32, n_nonzero, replace=False)
        # products spanning 30 orders of magnitude (1e10 and 1e-13 rate constants meet in the ZGB model)
        x[idx] = rng.randint(1, 400, n_nonzero) * 10.0 ** rng.uniform(-15, 15, n_nonzero)
        a, b = serial_prefix(x), packed_tier_prefix(x)
        assert a.tobytes() == b.tobytes(), (trial, x)


def test_packed_chain_edge_values():
    for x in (np.zeros(32), np.full(32, 1e-300), np.r_[np.zeros(31), 3.0], np.r_[1e308, 1e308, np.zeros(30)],
              np.r_[np.zeros(5), -0.0, 2.5, np.zeros(25)]):
        with np.errstate(over="ignore"):
            assert serial_prefix(x).tobytes() == packed_tier_prefix(x).tobytes()
