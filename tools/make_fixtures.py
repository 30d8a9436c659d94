#!/usr/bin/env python
"""Regenerate tests/golden/models/*.json from the kmos reference checkout.

Runs ONLY in the authoring container (needs /root/reference; ASE is replaced by tools/ase_shim,
JANAF tables by an ideal-gas stand-in).  For every model x backend the script

  1. builds the kmos ``Project`` (from the reference's .ini fixture or examples/render_*.py),
  2. calls the UNMODIFIED ``kmos.io.export_source`` to write the Fortran the reference would compile,
  3. parses that Fortran with ``kmos_b200.fortran_ir`` into the neutral rule-table IR and
  4. stores the IR (+ process rate expressions / parameters) as JSON.

The JSON files are the committed fixtures that the oracle, the CUDA engine, the tests and bench.py
load on the GPU box, where /root/reference does not exist.

usage: python tools/make_fixtures.py [--keep-fortran DIR]
"""
import argparse
import json
import os
import runpy
import shutil
import sys
import tempfile
import warnings
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("KMOS_REFERENCE", "/root/reference")

sys.path.insert(0, os.path.join(HERE, "ase_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

_janaf = MagicMock()
_janaf.__path__ = [tempfile.mkdtemp(prefix="janaf_stub_")]
sys.modules["janaf_data"] = _janaf
warnings.simplefilter("ignore")

import kmos.types  # noqa: E402
import kmos.io  # noqa: E402
import kmos.species  # noqa: E402

from kmos_b200 import fortran_ir  # noqa: E402
from kmos_b200.rates import standin_mu  # noqa: E402

# JANAF tables are not vendored (kmos/species.py:47-96 downloads them); use the documented stand-in
kmos.species.Species.mu = lambda self, T, p: standin_mu(self.name, T, p)

# lattice/species representations are ASE constructor strings used only by the viewer; store them verbatim
kmos.types.LayerList.__setattr__ = lambda self, key, value: self.__dict__.__setitem__(
    key, ("%s" % value) if key == "representation" else value)

MINI_101_INI = """[Meta]
author = Your Name
email = you@server.com
model_dimension = 2
model_name = fcc_100

[Species empty]
color = #FFFFFF

[Species CO]
representation = Atoms("CO", [[0, 0, 0], [0, 0, 1.17]])
color = #FF0000

[Lattice]
cell_size = 3.5 3.5 10.0

[Layer simple_cubic]
site hollow = (0.5, 0.5, 0.5)
color = #FFFFFF

[Parameter k_CO_ads]
value = 100
adjustable = True
min = 1
max = 1e13
scale = log

[Parameter k_CO_des]
value = 100
adjustable = True
min = 1
max = 1e13
scale = log

[Process CO_ads]
rate_constant = k_CO_ads
conditions = empty@hollow
actions = CO@hollow
tof_count = {'adsorption':1}

[Process CO_des]
rate_constant = k_CO_des
conditions = CO@hollow
actions = empty@hollow
tof_count = {'desorption':1}
"""


def project_from_ini(path_or_text):
    pt = kmos.types.Project()
    if os.path.exists(path_or_text):
        with open(path_or_text) as f:
            pt.import_ini_file(f)
    else:
        from io import StringIO
        pt.import_ini_file(StringIO(path_or_text))
    return pt


def project_dim(dim):
    """A small model of our own for the dimensions the reference's examples do not cover (they are all 2-d):
    one site per cell, species A / empty, adsorption, desorption and hops to every nearest neighbour along the
    `dim` axes (so the z and the 1-d index arithmetic of the kernels is exercised)."""
    from kmos.types import Project, Condition, Action
    import numpy as np
    pt = Project()
    pt.set_meta(author="kmos-b200 tests", email="none@example.org", model_name="hop%dd" % dim, model_dimension=dim)
    layer = pt.add_layer(name="sc")
    layer.add_site(name="a")
    pt.add_species(name="empty", color="#ffffff")
    pt.add_species(name="A", color="#ff0000", representation="Atoms('O')")
    pt.species_list.default_species = "empty"
    pt.add_parameter(name="k_ads", value=1.0, adjustable=True, min=0.1, max=10.0)
    pt.add_parameter(name="k_des", value=0.7, adjustable=True, min=0.1, max=10.0)
    pt.add_parameter(name="k_hop", value=2.0, adjustable=True, min=0.1, max=10.0)
    pt.lattice.cell = np.diag([1.0, 1.0, 1.0])
    center = pt.lattice.generate_coord("a.(0,0,0).sc")
    pt.add_process(name="ads", conditions=[Condition(species="empty", coord=center)],
                   actions=[Action(species="A", coord=center)], rate_constant="k_ads")
    pt.add_process(name="des", conditions=[Condition(species="A", coord=center)],
                   actions=[Action(species="empty", coord=center)], rate_constant="k_des")
    for axis in range(dim):
        for sign, tag in ((1, "p"), (-1, "m")):
            off = [0, 0, 0]
            off[axis] = sign
            nb = pt.lattice.generate_coord("a.(%d,%d,%d).sc" % tuple(off))
            pt.add_process(name="hop_%s%s" % ("xyz"[axis], tag),
                           conditions=[Condition(species="A", coord=center), Condition(species="empty", coord=nb)],
                           actions=[Action(species="empty", coord=center), Action(species="A", coord=nb)],
                           rate_constant="k_hop")
    return pt


def project_from_render_script(path, substitute=None):
    """Run an examples/render_*.py script, capturing the Project instead of saving XML.
    substitute: (old, new) text replacement applied to a temporary copy of the script (documented variants)."""
    if substitute:
        src = open(path).read()
        assert substitute[0] in src
        tmp_script = os.path.join(tempfile.mkdtemp(prefix="render_src_"), os.path.basename(path))
        with open(tmp_script, "w") as f:
            f.write(src.replace(substitute[0], substitute[1]))
        path = tmp_script
    captured = []
    orig_save = kmos.types.Project.save
    orig_action = kmos.types.ConditionAction.__init__

    def fake_save(self, *a, **k):
        captured.append(self)

    def tolerant_init(self, **kwargs):
        # examples/render_pairwise_interaction_otf.py:65 spells the keyword `cood=`; read it as `coord`
        if "cood" in kwargs:
            kwargs["coord"] = kwargs.pop("cood")
        orig_action(self, **kwargs)

    orig_process = kmos.types.Process.__init__

    def tolerant_process_init(self, **kwargs):
        # same example, line 72: Process(conditions=..., actions=...) instead of *_list
        for short, full in (("conditions", "condition_list"), ("actions", "action_list")):
            if short in kwargs:
                kwargs[full] = kwargs.pop(short)
        orig_process(self, **kwargs)

    kmos.types.Project.save = fake_save
    kmos.types.ConditionAction.__init__ = tolerant_init
    kmos.types.Process.__init__ = tolerant_process_init
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="render_")
    os.chdir(tmp)
    try:
        ns = runpy.run_path(path, run_name="__main__")
    finally:
        os.chdir(cwd)
        kmos.types.Project.save = orig_save
        kmos.types.ConditionAction.__init__ = orig_action
        kmos.types.Process.__init__ = orig_process
        shutil.rmtree(tmp, ignore_errors=True)
    if captured:
        return captured[-1]
    return ns["pt"]


MODELS = [
    # (fixture name, builder, backends)
    ("ab", lambda: project_from_ini(os.path.join(REF, "tests/test_run/AB_model.ini")),
     ["local_smart", "lat_int", "otf"]),
    ("mini_101", lambda: project_from_ini(MINI_101_INI), ["local_smart", "lat_int", "otf"]),
    ("zgb", lambda: project_from_render_script(os.path.join(REF, "examples/render_ZGB_model.py")),
     ["local_smart", "lat_int", "otf"]),
    ("ruo2", lambda: project_from_render_script(os.path.join(REF, "examples/render_co_oxidation_ruo2.py")),
     ["local_smart", "lat_int"]),
    ("pairwise", lambda: project_from_render_script(os.path.join(REF, "examples/render_pairwise_interaction.py")),
     ["lat_int", "local_smart"]),
    # the same example with three instead of two neighbour states (empty / CO / O): 3 + 3^4 = 84 processes,
    # a lat_int model beyond 64 processes (4 process segments per lane in the warp kernel)
    ("pairwise84", lambda: project_from_render_script(
        os.path.join(REF, "examples/render_pairwise_interaction.py"),
        substitute=("product(['empty', 'CO'], repeat=len(nn_coords))",
                    "product(['empty', 'CO', 'O'], repeat=len(nn_coords))")),
     ["lat_int", "local_smart"]),
    ("hop3d", lambda: project_dim(3), ["local_smart", "lat_int", "otf"]),
    ("hop1d", lambda: project_dim(1), ["local_smart", "lat_int"]),
    ("pairwise_otf",
     lambda: project_from_render_script(os.path.join(REF, "examples/render_pairwise_interaction_otf.py")),
     ["otf"]),
    # further examples of the reference (round 2): predator-prey on one site type with three species, a
    # five-species diffusion model, the sand-pile model, H on Pt(111) with two hollow sites per cell, and the
    # reference's own 1-d model
    ("lotka", lambda: project_from_render_script(os.path.join(REF, "examples/render_Lotka_Volterra_model.py")),
     ["local_smart", "lat_int"]),
    ("diffusion", lambda: project_from_render_script(os.path.join(REF, "examples/render_diffusion_model.py")),
     ["local_smart", "lat_int"]),
    ("sand", lambda: project_from_render_script(os.path.join(REF, "examples/render_sand_model.py")),
     ["local_smart", "lat_int"]),
    ("pt111", lambda: project_from_render_script(os.path.join(REF, "examples/render_Pt_111.py")),
     ["local_smart", "lat_int", "otf"]),
    ("einsd", lambda: project_from_render_script(os.path.join(REF, "examples/render_einsD.py")),
     ["local_smart", "lat_int", "otf"]),
    # multidentate adsorbates: species that occupy two and four sites at once
    ("multidentate", lambda: project_from_render_script(os.path.join(REF, "examples/multidentate.py")),
     ["local_smart", "lat_int", "otf"]),
]


def project_meta(pt):
    """Host-side inputs the hot path consumes but does not compute: rate expressions etc."""
    params = {}
    for p in pt.get_parameters():
        params[p.name] = {"value": p.value, "adjustable": bool(p.adjustable),
                          "min": p.min, "max": p.max, "scale": p.scale}
    procs = []
    for proc in pt.get_processes():
        procs.append({
            "name": proc.name,
            "rate_constant": proc.rate_constant,
            "otf_rate": getattr(proc, "otf_rate", None),
            "enabled": bool(proc.enabled),
            "tof_count": proc.tof_count if proc.tof_count else None,
            "conditions": [[c.species, c.coord.name, c.coord.layer, [int(x) for x in c.coord.offset]]
                           for c in proc.condition_list],
            "actions": [[a.species, a.coord.name, a.coord.layer, [int(x) for x in a.coord.offset]]
                        for a in proc.action_list],
            "bystanders": [[list(b.allowed_species), b.coord.name, b.coord.layer,
                            [int(x) for x in b.coord.offset], b.flag]
                           for b in getattr(proc, "bystander_list", [])],
        })
    return {"parameters": params, "process_defs": procs}


def copy_reference_goldens():
    """tests/test_run/_tmp_export_*/ref_procs_sites_*.log: the reference's only trajectory known-answer
    test (tests/test_run/test_run.py:40-70): AB model 20x20, seed 1, 10000 x (get_next_kmc_step,
    run_proc_nr).  Stored as int32[10000][2] (.npy) next to the model fixtures."""
    import ast
    import numpy as np
    outdir = os.path.join(REPO, "tests", "golden")
    arrays = []
    for backend in ("local_smart", "lat_int", "otf"):
        src = os.path.join(REF, "tests", "test_run", "_tmp_export_%s" % backend,
                           "ref_procs_sites_%s.log" % backend)
        with open(src) as f:
            arrays.append(np.asarray(ast.literal_eval(f.read()), dtype=np.int32))
    # the three logs are byte-identical upstream (test_run.py re-imports the cached local_smart
    # `kmc_model` extension for the 2nd and 3rd backend), so one copy is kept
    assert all(np.array_equal(arrays[0], a) for a in arrays[1:])
    out = os.path.join(outdir, "ab_ref_procs_sites.npy")
    np.save(out, arrays[0])
    print("wrote %s %s" % (out, arrays[0].shape))


# Fortran the reference itself keeps under version control (byte-golden exports of its own tests): parsed as is.
GOLDEN_EXPORTS = [
    ("pdopd", "local_smart", "tests/export_test/reference_pdopd_local_smart", "tests/export_test/pdopd.xml"),
    ("pdopd", "lat_int", "tests/export_test/reference_pdopd_lat_int", "tests/export_test/pdopd.xml"),
    # otf with 36 processes (the RuO2 test model) and with real bystanders (interacting ZGB)
    ("ruo2default", "otf", "tests/export_test/reference_export_otf", "tests/export_test/default.xml"),
    ("intzgb", "otf", "tests/export_test/reference_export_intZGB_otf", "tests/export_test/intZGB_otf.xml"),
]


def xml_parameters(path):
    """<parameter name= value= adjustable= min= max= scale=/> entries of a kmos project XML (plain ElementTree:
    the otf rate tables need the user parameters, which the exported Fortran receives only at run time)."""
    import xml.etree.ElementTree as ET
    out = {}
    for p in ET.parse(path).getroot().iter("parameter"):
        a = p.attrib
        out[a["name"]] = {"value": a["value"], "adjustable": a.get("adjustable") == "True",
                          "min": float(a.get("min", 0.0)), "max": float(a.get("max", 0.0)),
                          "scale": a.get("scale", "linear")}
    return out


def golden_export_fixtures(outdir):
    for name, backend, rel, xml in GOLDEN_EXPORTS:
        ir = fortran_ir.parse_export_dir(os.path.join(REF, rel), backend)
        ir["fixture"] = {"model": name, "backend": backend, "settings_written": False,
                         "generator": "%s (reference, committed Fortran) -> kmos_b200.fortran_ir" % rel}
        ir.update({"parameters": xml_parameters(os.path.join(REF, xml)), "process_defs": []})
        out = os.path.join(outdir, "%s_%s.json" % (name, backend))
        with open(out, "w") as f:
            json.dump(ir, f, separators=(",", ":"), sort_keys=True)
        print("wrote %s (%d procs, %d bytes)" % (out, len(ir["procs"]), os.path.getsize(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--keep-fortran", default=None, help="directory to keep the generated Fortran in")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()

    outdir = os.path.join(REPO, "tests", "golden", "models")
    os.makedirs(outdir, exist_ok=True)
    if not args.only:
        copy_reference_goldens()
    if not args.only or args.only == "pdopd":
        golden_export_fixtures(outdir)
    if args.only == "pdopd":
        return
    for name, builder, backends in MODELS:
        if args.only and args.only != name:
            continue
        for backend in backends:
            pt = builder()
            fdir = (os.path.join(args.keep_fortran, "%s_%s" % (name, backend)) if args.keep_fortran
                    else tempfile.mkdtemp(prefix="kmos_export_"))
            if os.path.exists(fdir):
                shutil.rmtree(fdir)
            settings_ok = True
            try:
                kmos.io.export_source(pt, fdir, code_generator=backend)
            except Exception as e:  # write_settings is the last step; Fortran is complete by then
                settings_ok = False
                print("  [%s/%s] export_source raised after writing Fortran: %s" %
                      (name, backend, str(e).splitlines()[0][:100]))
            ir = fortran_ir.parse_export_dir(fdir, backend)
            ir["fixture"] = {"model": name, "backend": backend, "settings_written": settings_ok,
                             "generator": "kmos.io.export_source (reference, unmodified) -> kmos_b200.fortran_ir"}
            ir.update(project_meta(pt))
            out = os.path.join(outdir, "%s_%s.json" % (name, backend))
            with open(out, "w") as f:
                json.dump(ir, f, separators=(",", ":"), sort_keys=True)
            print("wrote %s (%d procs, %d bytes)" % (out, len(ir["procs"]), os.path.getsize(out)))
            if not args.keep_fortran:
                shutil.rmtree(fdir, ignore_errors=True)


if __name__ == "__main__":
    main()
