"""Exporter hook: model_tables.{json,bin} written next to generated Fortran (authoring container only)."""
import os
import shutil

import numpy as np
import pytest

from kmos_b200 import export, tables

REF = os.environ.get("KMOS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "tests", "export_test", "reference_export")


@pytest.mark.skipif(not os.path.isdir(SRC), reason="reference checkout absent")
def test_export_tables_next_to_fortran(tmp_path):
    d = tmp_path / "src"
    shutil.copytree(SRC, d)
    ir, blob, info = export.export_tables(str(d))
    assert ir["backend"] == "local_smart" and len(ir["procs"]) == 36
    on_disk = np.fromfile(d / "model_tables.bin", dtype=np.int32)
    assert np.array_equal(on_disk, blob) and on_disk[0] == tables.MAGIC
    assert tables.load_ir(str(d / "model_tables.json"))["procs"] == ir["procs"]
    # the image is accepted by the C-ABI (no GPU needed for model_create)
    import ctypes
    from kmos_b200 import capi
    capi.build()
    h = ctypes.c_void_p()
    assert capi.lib().kmos_b200_model_create(on_disk, on_disk.size, ctypes.byref(h)) == 0
    capi.lib().kmos_b200_model_destroy(h)
    # ... and the exporter emitted the model's CUDA proclist next to the Fortran (north_star)
    cu = d / "proclist_model.cu" if (d / "proclist_model.cu").exists() else info["proclist_cu"]
    text = open(cu).read()
    assert "KB_GEN_MODULE(KbModel, kb_info)" in text and "static constexpr int P = 36" in text
    assert "static constexpr int LPR = 16" in text          # RuO2: 16 lanes per replica, two replicas per warp
    assert text.count("// process ") == 36 and "replace_species(site + (0,0,0) type" in text
    # lattice geometry for the front-end
    assert ir["unit_cell_size"][0][0] == 10.0 and len(ir["site_positions"]) == 2
