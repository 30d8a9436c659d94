"""KMC_Model mirror on the GPU: the reference-shaped outputs (TOFs, coverages, std header/data, config
round trip) computed from the batched engine must equal the same quantities computed from the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from kmos_b200 import rates as rates_mod, tables

pytestmark = pytest.mark.gpu


def _oracle_for(ir, size, seed, replica, overrides):
    from oracle import oracle
    blob, info = tables.build_blob(ir, with_device=False)
    return oracle.Oracle(blob, size, seed=seed, replica=replica, rates=rates_mod.model_rates(ir, overrides))


def test_std_header_and_sampled_data_match_oracle():
    from kmos_b200.model import KMC_Model
    path = os.path.join(GOLDEN, "models", "ab_local_smart.json")
    ir = tables.load_ir(path)
    points = [{"p_COgas": 1.0, "p_O2gas": 1.0}, {"p_COgas": 0.3, "p_O2gas": 2.0}, {"p_COgas": 3.0, "p_O2gas": 0.5}]
    with KMC_Model(path, size=[12, 10], n_replicas=3, parameters=points, random_seed=5) as m:
        hdr = m.get_std_header()
        assert hdr == "#p_COgas p_O2gas TOF A_default_a B_default_a empty_default_a kmc_time simulated_time kmc_steps\n"
        assert m.get_backend() == "local_smart"
        m.do_steps(2000)
        rows = m.get_std_sampled_data_all(samples=4, sample_size=4000, tof_method="procrates")
        rows_i = None
        for r, pt in enumerate(points):
            o = _oracle_for(ir, [12, 10], 5 + r, r, pt)
            o.do_steps(2000)
            t0, ps0 = o.kmc_time, o.procstat.copy()
            occs, tofs, dts, sts = [], [], [], []
            t_prev, ps_prev = t0, ps0
            for _ in range(4):
                o.do_steps(1000)
                dt = o.kmc_time - t_prev
                tofs.append(m.tof_matrix @ ((o.procstat - ps_prev) / dt / 120.0))
                occs.append(o.occupation.flatten())
                dts.append(dt)
                sts.append(o.L.kmos_oracle_kmc_time_step(o.h))
                t_prev, ps_prev = o.kmc_time, o.procstat.copy()
            tof_mean = np.average(tofs, axis=0, weights=dts)
            occ_mean = np.average(occs, axis=0, weights=sts)
            ref = [pt["p_COgas"], pt["p_O2gas"]] + list(tof_mean) + list(occ_mean) + [o.kmc_time - t0, o.kmc_time, 4000]
            np.testing.assert_allclose(rows[r], ref, rtol=1e-9, atol=1e-300)
        s = m.get_std_sampled_data(2, 1000, tof_method="integ", output="str", replica=1)
        assert len(s.split()) == len(hdr[1:].split())
        d = m.get_std_sampled_data(2, 1000, output="dict")
        assert set(d) == set(hdr[1:].split())


def test_configuration_roundtrip_put_and_f2py_shim(tmp_path):
    from kmos_b200.model import KMC_Model
    from oracle import oracle
    path = os.path.join(GOLDEN, "models", "ruo2_local_smart.json")
    ir = tables.load_ir(path)
    with KMC_Model(path, size=[6, 5], n_replicas=2, random_seed=3) as m:
        m.do_steps(500)
        cfg = m._get_configuration(replica=1)
        assert cfg.shape == (6, 5, 1, 2) and cfg.dtype == np.int8
        m.dump_config(str(tmp_path / "cfg"), replica=1)
        m.load_config(str(tmp_path / "cfg"), replica=0)
        assert np.array_equal(m._get_configuration(0), cfg)
        # put(): species by name, book-keeping re-adjusted like KMC_Model._adjust_database
        m.put([2, 3, 0, 1], "CO", replica=0)
        assert m._get_configuration(0)[2, 3, 0, 0] == ir["species"].index("CO")
        assert m.lattice.get_species([2, 3, 0, 1]) == ir["species"].index("CO")
        # avail_sites after put must equal the oracle's after the same _set_configuration
        o = _oracle_for(ir, [6, 5], 3, 0, {})
        o.do_steps(0)
        # replay: oracle replica 1 to get cfg, then replica 0 path
        o1 = _oracle_for(ir, [6, 5], 4, 1, {})
        o1.do_steps(500)
        o0 = _oracle_for(ir, [6, 5], 3, 0, {})
        o0.do_steps(500)
        assert o0.set_configuration(o1.lattice) == 0
        lat = o0.lattice.copy()
        lat[m.lattice.calculate_lattice2nr([2, 3, 0, 1]) - 1] = ir["species"].index("CO")
        assert o0.set_configuration(lat) == 0
        assert np.array_equal(m.batch.avail_sites(0), o0.avail_sites)
        assert np.array_equal(m.batch.nr_of_sites[0], o0.nr_of_sites)
        # f2py-shaped accessors
        assert m.base.get_kmc_step() == 500 and m.proclist.nr_of_proc == 36
        assert m.base.get_nrofsites(1) == int(o0.nr_of_sites[0])
        assert abs(m.base.get_accum_rate(36) - o0_total(o0)) <= 1e-12 * o0_total(o0)
        assert m.proclist.co_adsorption_cus == 1 and m.proclist.empty == ir["species"].index("empty")


def o0_total(o):
    o.L.kmos_oracle_update_accum_rate(o.h)
    return float(o.accum_rates[-1])


def test_model_runner_writes_reference_shaped_dat(tmp_path):
    from kmos_b200.runner import ModelRunner, PressureParameter, TemperatureParameter

    class Scan(ModelRunner):
        T = TemperatureParameter(600)
        p_COgas = PressureParameter(min=0.5, max=2.0, steps=3)
        p_O2gas = PressureParameter(1)

    path = os.path.join(GOLDEN, "models", "ab_local_smart.json")
    out = str(tmp_path / "Scan.dat")
    runner = Scan(path, size=[10, 10], seeds=2)
    assert [p["p_COgas"] for p in runner.grid_points()] == pytest.approx([0.5, 1.0, 2.0])
    header, rows = runner.run(init_steps=2000, sample_steps=4000, samples=2, outfile=out)
    assert rows.shape == (3, len(header[1:].split()))
    lines = open(out).read().splitlines()
    assert lines[0] == header.strip()
    data = np.loadtxt(out)
    assert data.shape == rows.shape
    cols = header[1:].split()
    np.testing.assert_allclose(data[:, cols.index("p_COgas")], [0.5, 1.0, 2.0], rtol=1e-5)
    assert np.all(data[:, cols.index("kmc_steps")] == 4000)
    # coverages of one site type sum to one
    occ = [c for c in cols if c.endswith("_default_a")]
    np.testing.assert_allclose(data[:, [cols.index(c) for c in occ]].sum(axis=1), 1.0, rtol=1e-4)
