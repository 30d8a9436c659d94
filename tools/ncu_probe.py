#!/usr/bin/env python
"""Small workload for `ncu --set full`: one model, two launches of the step kernel (profile the second).

    python tools/ncu_probe.py [model] [R] [n] [LxL] [auto|generated|smem|warp_hbm]
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from kmos_b200 import capi, engine, tables, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ruo2_local_smart"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 444
n = int(sys.argv[3]) if len(sys.argv) > 3 else 300
size = [int(x) for x in sys.argv[4].split("x")] if len(sys.argv) > 4 else [20, 20]
kern = sys.argv[5] if len(sys.argv) > 5 else "auto"
ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", name + ".json"))
m = engine.Model(ir=ir)
kw = {}
if kern == "generated":
    kw = dict(proclist="build", kernel=capi.KERNEL_GENERATED)
elif kern != "auto":
    kw = dict(kernel={"smem": capi.KERNEL_SMEM, "warp_hbm": capi.KERNEL_WARP_HBM, "generic": capi.KERNEL_GENERIC}[kern])
b = engine.Batch(m, R, size, rates=workloads.rates_for(name, ir, R), **kw)
b.do_steps(n)   # warm-up launch (skipped by ncu -s 1 of the step kernel)
b.do_steps(n)
b.synchronize()
print(b.kernel_info(), int((b.status == 0).sum()))
