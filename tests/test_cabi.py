"""The C-ABI library loads and exports every symbol include/kmos_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REPO, load_model
from kmos_b200 import capi


@pytest.fixture(scope="module")
def built():
    capi.build()
    return capi.LIB


def test_header_symbols_are_exported(built):
    header = open(os.path.join(REPO, "include", "kmos_b200.h")).read()
    declared = set(re.findall(r"\b(kmos_b200_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(capi.EXPORTED), declared ^ set(capi.EXPORTED)
    L = ctypes.CDLL(built)
    for name in sorted(declared):
        assert hasattr(L, name), "%s not exported" % name


def test_model_create_validates_blob(built):
    L = capi.lib()
    ir, blob, info = load_model("ruo2_local_smart")
    h = ctypes.c_void_p()
    assert L.kmos_b200_model_create(blob, blob.size, ctypes.byref(h)) == 0
    assert L.kmos_b200_model_nproc(h) == 36
    assert L.kmos_b200_model_nspecies(h) == 3
    assert L.kmos_b200_model_spuck(h) == 2
    L.kmos_b200_model_destroy(h)
    bad = blob.copy()
    bad[0] = 0
    assert L.kmos_b200_model_create(bad, bad.size, ctypes.byref(h)) == -2
    assert b"KB20" in L.kmos_b200_last_error()


def test_no_cpu_fallback(built):
    """Without a CUDA device batch creation must fail loudly instead of computing on the host."""
    L = capi.lib()
    if L.kmos_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    ir, blob, info = load_model("mini_101_local_smart")
    from kmos_b200 import engine
    model = engine.Model(ir=ir, blob=blob, info=info)
    with pytest.raises(capi.KmosB200Error, match="no CUDA device"):
        engine.Batch(model, 4, [4, 4])
    with pytest.raises(capi.KmosB200Error, match="no CUDA device"):
        engine.Fleet(model, 4, [4, 4], gpu_ids=[0, 1])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "kmos_b200")
    for root, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "kmos_oracle" not in text \
                    or f == "kb_common.h", f


def test_philox_hook_matches_oracle_cpu(built):
    from oracle import oracle
    L = capi.lib()
    for seed, rep, step in [(1, 0, 0), (2**40 + 17, 123, 2**33 + 5)]:
        ref = oracle.philox_step(seed, rep, step)
        assert list(ref) == [L.kmos_b200_philox_next(seed, rep, step, s) for s in range(3)]


def test_fortran_interface_block_matches_the_header():
    """INTEGRATION.md section 3 (ISO_C_BINDING interface for the template-side proclist): every bind(C) name must be
    an exported symbol and take as many arguments as the C declaration (no Fortran compiler here to check more)."""
    import re
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    block = text[text.index("module kmos_b200_c"):text.index("end module")]
    header = open(os.path.join(REPO, "include", "kmos_b200.h")).read()
    funcs = re.findall(r"function\s+(\w+)\s*\(([^)]*)\)\s*&?\s*bind\(C,\s*name=\"(\w+)\"\)", block, re.S)
    assert len(funcs) >= 6
    for fname, fargs, cname in funcs:
        assert fname == cname and cname in capi.EXPORTED, cname
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % cname, header, re.S)
        assert m, "not declared in the header: %s" % cname
        n_c = 0 if m.group(1).strip() in ("", "void") else m.group(1).count(",") + 1
        n_f = len([a for a in fargs.replace("&", "").split(",") if a.strip()])
        assert n_c == n_f, (cname, n_c, n_f)


def test_fleet_create_validates_its_arguments_before_touching_cuda(built):
    """SURVEY 8b gpu_ids[]: an empty device list or no replicas is an argument error (-1) with or without a GPU."""
    L = capi.lib()
    ir, blob, info = load_model("mini_101_local_smart")
    m, f = ctypes.c_void_p(), ctypes.c_void_p()
    assert L.kmos_b200_model_create(blob, blob.size, ctypes.byref(m)) == 0
    size = np.array([4, 4, 1], dtype=np.int32)
    ids = np.array([0], dtype=np.int32)
    assert L.kmos_b200_fleet_create(m, 4, size, None, ids, 0, ctypes.byref(f)) == -1
    assert L.kmos_b200_fleet_create(m, 0, size, None, ids, 1, ctypes.byref(f)) == -1
    assert b"fleet_create" in L.kmos_b200_last_error()
    L.kmos_b200_fleet_destroy(None)  # a no-op, like free(NULL)
    L.kmos_b200_model_destroy(m)
