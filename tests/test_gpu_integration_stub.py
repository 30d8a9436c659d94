"""INTEGRATION.md section 2 shows the ctypes stub a kmos maintainer would drop into kmos/run/__init__.py:86 in place
of `from kmc_model import base, lattice, proclist`.  This test executes that very text (extracted from the
document) against the built library, so the documentation cannot drift from the C-ABI."""
import os
import re

import numpy as np
import pytest

from conftest import REPO, load_model
from kmos_b200 import capi

pytestmark = pytest.mark.gpu


def test_documented_ctypes_stub_runs(tmp_path, monkeypatch):
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    section = text[text.index("## 2. Python"):text.index("## 3. Fortran")]
    code = re.search(r"```python\n(.*?)```", section, re.S).group(1)
    code = code.replace('ctypes.CDLL("libkmos_b200.so")', "ctypes.CDLL(%r)" % capi.LIB)
    ir, blob, info = load_model("ab_local_smart")
    blob.tofile(str(tmp_path / "model_tables.bin"))
    monkeypatch.chdir(tmp_path)
    ns = {}
    exec(compile(code, "INTEGRATION.md#2", "exec"), ns)
    proclist, base = ns["proclist"], ns["base"]
    proclist.init([7, 6], "stub", ir["default_layer"], 42, True)
    for i in range(len(ir["procs"])):
        base.set_rate_const(i + 1, 1.0 + 0.1 * i)
    proc, site = proclist.get_next_kmc_step()
    assert 1 <= proc <= len(ir["procs"]) and 1 <= site <= 42
    proclist.run_proc_nr(proc, site)
    proclist.do_kmc_steps(500)
    t = base.get_kmc_time()
    assert t > 0 and np.isfinite(t)
    # same trajectory as the oracle under the same seed / rates
    from oracle import oracle
    o = oracle.Oracle(blob, [7, 6], seed=42, replica=0, rates=np.array([1.0 + 0.1 * i for i in range(len(ir["procs"]))]))
    p2, s2, _st = o.get_next_kmc_step()
    assert (p2, s2) == (proc, site)
    o.run_proc_nr(p2, s2)
    o.do_steps(500)
    assert abs(t - o.kmc_time) <= 1e-12 * o.kmc_time
