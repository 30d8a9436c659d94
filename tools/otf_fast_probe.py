#!/usr/bin/env python
"""Config E on the otf production kernel (kb_otf_fast.cuh): timing, or a short workload for ncu.

    python tools/otf_fast_probe.py [R] [n] [LxL] [time|ncu]
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from kmos_b200 import capi, engine, otf, tables, workloads  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 3552
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
size = [int(x) for x in sys.argv[3].split("x")] if len(sys.argv) > 3 else [256, 256]
mode = sys.argv[4] if len(sys.argv) > 4 else "time"
name = os.environ.get("KMOS_B200_PROBE_MODEL", "pairwise_otf_otf")
ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", name + ".json"))
m = engine.Model(ir=ir)
rates = workloads.rates_for("pairwise" if name.startswith("pairwise") else "other", ir, R)
lut = np.tile(otf.build_lut(ir, m.info, rates[0]), (R, 1))
b = engine.Batch(m, R, size, rates=rates, lut=lut)
b.do_steps(10)
b.select_kernel(capi.KERNEL_OTF_FAST)
b.do_steps(n // 4 if mode == "time" else n)
b.synchronize()
if mode == "time":
    ts = []
    for _ in range(3):
        b.timer_start()
        b.do_steps(n)
        ts.append(b.timer_stop())
    print("otf_fast %s R=%d: %.3e kMC steps/s (%.2f ms per %d steps) ok=%d  procstat %s" %
          (size, R, R * n / (np.mean(ts) * 1e-3), np.mean(ts), n, int((b.status == 0).sum()),
           b.procstat.sum(axis=0).tolist()))
else:
    b.do_steps(n)
    b.synchronize()
    print(b.kernel_info())
