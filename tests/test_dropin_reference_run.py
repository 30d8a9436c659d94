"""The unmodified reference front-end on top of the kmc_model drop-in (kmos_b200/dropin/kmc_model).

``import kmos.run`` (the reference's own file, untouched) resolves ``from kmc_model import base, lattice,
proclist`` to our package and ``import kmc_settings`` to the file the reference's exporter writes.  The
container has the reference but no GPU, so the batch behind the package is the oracle-backed test double
(tests/oracle_engine.py); the same package is exercised on CUDA by tests/test_gpu_dropin.py.

With the double in gfortran-RNG mode the loop of the reference's own tests/test_run/test_run.py:46-53
(``get_next_kmc_step`` / ``run_proc_nr`` x 10 000) must reproduce the reference's golden log byte for byte.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "kmos")), reason="reference checkout not present")

DRIVER = r'''
import json, os, sys, tempfile, warnings
from unittest.mock import MagicMock
repo, ref, export_dir, backend, mode = sys.argv[1:6]
warnings.simplefilter("ignore")
janaf = MagicMock(); janaf.__path__ = [tempfile.mkdtemp(prefix="janaf_stub_")]
sys.modules["janaf_data"] = janaf                      # as the reference's own tests/conftest.py does
sys.path[:0] = [os.path.join(repo, "kmos_b200", "dropin"), export_dir, os.path.join(repo, "tools", "ase_shim"),
                ref, repo, os.path.join(repo, "tests")]
import numpy as np
import kmos.types, kmos.io
from kmos_b200 import export
if mode == "bystanders":   # examples/render_pairwise_interaction_otf.py: a desorption rate that depends on nr_CO_1nn
    sys.path.insert(0, os.path.join(repo, "tools"))
    import make_fixtures
    pt = make_fixtures.project_from_render_script(os.path.join(ref, "examples", "render_pairwise_interaction_otf.py"))
else:
    pt = kmos.types.Project()
    with open(os.path.join(ref, "tests", "test_run", "AB_model.ini")) as f:
        pt.import_ini_file(f)
export.export_source(pt, export_dir, code_generator=backend)   # the reference's exporter + model_tables.json
os.chdir(export_dir)

import kmc_model                                          # our drop-in package
from kmc_model import _runtime
import oracle_engine
from oracle import oracle
rng = oracle.RNG_GFORTRAN if mode == "golden" else oracle.RNG_PHILOX
_runtime.batch_factory = lambda ir, size, seed, layer: oracle_engine.OracleBatch(ir, size, seed, layer, rng=rng)

import kmos.run                                           # UNMODIFIED reference front-end
assert kmos.run.base is kmc_model.base and kmos.run.settings is not None
out = {}
with kmos.run.KMC_Model(print_rates=False, banner=False, size=(8 if mode == "bystanders" else None)) as model:
    out["size"] = [int(x) for x in model.size]
    out["backend"] = model.get_backend()
    if mode == "bystanders":
        # proclist_pars.byst_<proc> / rate_<proc>(nr_vars) behind KMC_Model.rate_constants (run/__init__.py:1851-1899)
        rc = model.rate_constants
        out["byst"] = rc.bystanders(pattern="CO_desorption", interactive=False)
        out["rate0"] = rc._rate("CO_desorption")
        out["rate2"] = rc._rate("CO_desorption", nr_CO_1nn=2)
        out["base_rate"] = float(model.base.get_rate(model.proclist.co_desorption))
        model.do_steps(3000)
        # the live rates_matrix holds rate_<proc>(environment) for every registered (proc, site)
        b = _runtime.batch()
        p = int(model.proclist.co_desorption)
        row = b.o.rates_matrix_row(p)
        n = int(b.o.nr_of_sites[p - 1])
        vals = sorted(set(round(float(v), 6) for v in row[:n]))
        table = sorted(set(round(rc._rate("CO_desorption", nr_CO_1nn=k), 6) for k in range(5)))
        out["row_in_table"] = all(v in table for v in vals) and n > 0
        # a parameter change goes through update_user_parameter + recalculate_rates_matrix
        before = rc._rate("CO_desorption", nr_CO_1nn=1)
        model.parameters.E_CO_nn = 0.05
        out["rate_changed"] = rc._rate("CO_desorption", nr_CO_1nn=1) != before
        model.do_steps(2000)
        out["kmc_step"] = int(model.base.get_kmc_step())
        row = b.o.rates_matrix_row(p)
        n = int(b.o.nr_of_sites[p - 1])
        table = sorted(set(round(rc._rate("CO_desorption", nr_CO_1nn=k), 6) for k in range(5)))
        out["row_in_new_table"] = all(round(float(v), 6) in table for v in row[:n]) and n > 0
    elif mode == "golden":
        procs_sites = []
        for i in range(10000):
            proc, site = model.get_next_kmc_step()
            procs_sites.append((int(proc.real), int(site.real)))
            model.run_proc_nr(proc, site)
        out["procs_sites"] = procs_sites
    else:
        model.do_steps(10000)
        atoms = model.get_atoms(geometry=False)
        out["kmc_step"] = int(atoms.kmc_step)
        out["kmc_time"] = float(atoms.kmc_time)
        out["occupation"] = np.asarray(atoms.occupation).tolist()
        out["procstat"] = [int(x) for x in atoms.procstat]
        out["header"] = model.get_std_header()
        out["row"] = model.get_std_sampled_data(samples=2, sample_size=2000, tof_method="integ")
        # put(): Python-side replace_species + the full _adjust_database pass, flushed to the engine in one call
        s0 = int(model.lattice.get_species([1, 2, 0, 1]))
        new = (s0 + 1) % 3
        model.put([1, 2, 0, 1], new)
        out["put_ok"] = int(model.lattice.get_species([1, 2, 0, 1])) == new
        model.do_steps(500)
        cfg = model._get_configuration()
        model._set_configuration(cfg)
        model.do_steps(500)
        out["kmc_step_end"] = int(model.base.get_kmc_step())
        out["avail_ok"] = model.base.get_avail_site(1, 1, 1) >= 0
        # the reporting and editing helpers of the front-end, all on the f2py-shaped getters
        out["coverages"] = model.print_coverages(to_stdout=False)
        out["procstat_txt"] = model.print_procstat(to_stdout=False)
        out["kmc_state"] = model.print_kmc_state(to_stdout=False)
        out["accum"] = model.print_accum_rate_summation(to_stdout=False)
        out["avail_proc1"] = len(model.get_avail(1))
        out["nr_of_sites1"] = int(model.base.get_nrofsites(1))
        out["nr2site"] = [str(x) for x in model.nr2site(5)]
        name0 = model.rate_constants.names()[0]
        model.rate_constants.set(name0, 12.5)
        out["rate_set"] = float(model.base.get_rate(1))   # set() addresses processes by sorted position
        model.parameters.T = 550
        model.do_steps(300)
        model.dump_config("cfg_test")
        lat_before = np.array(model._get_configuration())
        step_before = int(model.base.get_kmc_step())
        model.do_steps(300)
        model.load_config("cfg_test")
        out["load_config_ok"] = bool(np.array_equal(np.array(model._get_configuration()), lat_before))
        model.double()
        out["size_doubled"] = [int(x) for x in model.size]
        tiled = np.array(model._get_configuration())
        out["double_tiles"] = bool(np.array_equal(tiled[:lat_before.shape[0], :lat_before.shape[1]], lat_before) and
                                   np.array_equal(tiled[lat_before.shape[0]:, lat_before.shape[1]:], lat_before))
        model.do_steps(300)
        out["kmc_step_doubled"] = int(model.base.get_kmc_step())
        model.reset()
        out["kmc_step_reset"] = int(model.base.get_kmc_step())
print("RESULT " + json.dumps(out))
'''


def _run(tmp_path, backend, mode):
    script = tmp_path / "driver.py"
    script.write_text(DRIVER)
    export_dir = tmp_path / ("export_" + backend)
    export_dir.mkdir()
    p = subprocess.run([sys.executable, str(script), REPO, REF, str(export_dir), backend, mode],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    import json
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_reference_test_run_loop_reproduces_the_golden_log(tmp_path):
    out = _run(tmp_path, "local_smart", "golden")
    ref = np.load(os.path.join(HERE, "golden", "ab_ref_procs_sites.npy"))
    assert out["size"] == [20, 20] and out["backend"] == "local_smart"
    assert np.array_equal(np.asarray(out["procs_sites"]), ref)


def test_otf_bystander_rates_through_the_unmodified_front_end(tmp_path):
    """proclist_pars of an otf model with a bystander-dependent rate: byst_<proc>, rate_<proc>(nr_vars),
    update_user_parameter and recalculate_rates_matrix as kmos.run uses them."""
    out = _run(tmp_path, "otf", "bystanders")
    assert out["backend"] == "otf" and out["size"] == [8, 8]
    assert "nr_CO_1nn" in out["byst"]
    assert out["rate0"] == pytest.approx(out["base_rate"], rel=1e-12) and out["rate2"] != out["rate0"]
    assert out["row_in_table"] and out["rate_changed"] and out["row_in_new_table"]
    assert out["kmc_step"] == 5000


@pytest.mark.parametrize("backend", ["local_smart", "lat_int", "otf"])
def test_unmodified_kmc_model_front_end_runs_on_the_dropin(tmp_path, backend):
    out = _run(tmp_path, backend, "api")
    assert out["backend"] == backend
    assert out["kmc_step"] == 10000 and out["kmc_time"] > 0
    assert sum(out["procstat"]) == 10000
    assert abs(sum(sum(r) for r in out["occupation"]) - 1.0) < 1e-12
    assert out["header"].startswith("#") and out["header"].rstrip().endswith("kmc_time simulated_time kmc_steps")
    assert len(out["row"].split()) == len(out["header"].split())
    assert out["put_ok"] and out["avail_ok"] and out["kmc_step_end"] == 10000 + 2000 + 1000
    assert "default_a" in out["coverages"] and "A_adsorption" in out["procstat_txt"]
    assert "kmc steps" in out["kmc_state"] and "A_adsorption" in out["accum"]
    assert out["avail_proc1"] == out["nr_of_sites1"] and out["nr2site"][3] == "default_a"
    assert out["rate_set"] == 12.5
    assert out["load_config_ok"]
    assert out["size_doubled"] == [40, 40] and out["double_tiles"]
    assert out["kmc_step_doubled"] >= 300 and out["kmc_step_reset"] == 0
