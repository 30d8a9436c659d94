"""The generated-Fortran parser on (a) a hand-written miniature of the generator's grammar and (b) the
reference's own byte-golden exports (tests/export_test/reference_*), when the reference checkout is present
(authoring container only; skipped on the GPU box)."""
import os
import textwrap

import numpy as np
import pytest

from kmos_b200 import fortran_ir, tables

REF = os.environ.get("KMOS_REFERENCE", "/root/reference")


def test_logical_lines_join_continuations_and_strip_comments():
    text = textwrap.dedent("""
        ! a comment
        integer(kind=iint), parameter, public :: co = 0   ! trailing
        case(a, b,&
         c&
        )
        x = 0; return
    """)
    lines = fortran_ir._logical_lines(text)
    assert lines == ["integer(kind=iint), parameter, public :: co = 0", "case(a, b, c )", "x = 0", "return"]


def test_fixture_statement_kinds():
    """Every committed fixture only contains statement kinds the byte-code knows."""
    gold = os.path.join(os.path.dirname(__file__), "golden", "models")
    kinds = set()

    def walk(block):
        for st in block:
            kinds.add(st[0])
            if st[0] == "select":
                for _k, b in st[2]:
                    walk(b)
            elif st[0] == "if_can":
                walk(st[3])
    for f in sorted(os.listdir(gold)):
        ir = tables.load_ir(os.path.join(gold, f))
        for block in ir["routines"].values():
            walk(block)
        for block in ir["nli"].values():
            walk(block)
        for g in ir["gr"].values():
            walk(g["body"])
        blob, info = tables.build_blob(ir)
        assert blob[0] == tables.MAGIC and blob[1] == tables.VERSION
        assert blob.dtype == np.int32
    assert kinds <= {"replace", "if_can", "del", "add", "update_rate", "select", "del_all", "call", "return", "inc"}


GOLDEN_EXPORTS = [
    ("reference_export", "local_smart", 36),
    ("reference_export_lat_int", "lat_int", 36),
    ("reference_export_otf", "otf", 36),
    ("reference_export_intZGB_otf", "otf", 10),
]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests", "export_test")), reason="reference checkout absent")
@pytest.mark.parametrize("dirname,backend,nproc", GOLDEN_EXPORTS)
def test_parses_reference_golden_exports(dirname, backend, nproc):
    path = os.path.join(REF, "tests", "export_test", dirname)
    ir = fortran_ir.parse_export_dir(path, backend)
    assert len(ir["procs"]) == nproc and ir["backend"] == backend
    blob, info = tables.build_blob(ir)
    assert blob[4] == nproc
    if backend == "local_smart":
        assert info["device"]["supported"], info["device"]
    if backend == "otf":
        assert info["lut_total"] >= nproc


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests", "export_test")), reason="reference checkout absent")
@pytest.mark.parametrize("dirname,backend", [("reference_pdopd_local_smart", "local_smart"),
                                             ("reference_pdopd_lat_int", "lat_int")])
def test_parses_multilattice_exports(dirname, backend):
    """Pd/PdO: create_<site>(site, species) / annihilate_<site>(site, species) (kmos/io/__init__.py:2445-2560)
    become one routine instance per species passed; the declared null_species is what fills absent sites."""
    ir = fortran_ir.parse_export_dir(os.path.join(REF, "tests", "export_test", dirname), backend)
    assert len(ir["procs"]) == 46 and ir["spuck"] == 25 and len(ir["layers"]) == 2
    assert ir["null_species"] == ir["species"].index("null_species") == 4
    if backend == "local_smart":
        inst = [n for n in ir["routines"] if "@" in n]
        assert any(n.startswith("create_") for n in inst) and any(n.startswith("annihilate_") for n in inst)
        st = ir["routines"]["create_Pd100_h1@2"][0]
        assert st[0] == "replace" and st[2:] == [4, 2]  # replace_species(site, null_species, species)
    blob, info = tables.build_blob(ir)
    assert blob[4] == 46 and (blob[7] >> 16) - 1 == 4
    committed = tables.load_ir(os.path.join(os.path.dirname(__file__), "golden", "models",
                                            "pdopd_%s.json" % backend))
    assert committed["routines"] == ir["routines"] and committed["run_proc"] == ir["run_proc"]
