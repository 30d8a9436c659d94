#!/bin/bash
# how much would more resident replicas buy?  smaller lattices hold more replicas per SM
for S in 20 18 16 14; do
python - <<PY
import sys
sys.path.insert(0, ".")
from kmos_b200 import capi, engine, tables, workloads
ir = tables.load_ir("tests/golden/models/ruo2_local_smart.json")
m = engine.Model(ir=ir)
R, n = 16384, 5000
b = engine.Batch(m, R, [$S, $S], rates=workloads.rates_for("ruo2", ir, R), kernel=capi.KERNEL_GENERATED)
info = b.kernel_info()
b.do_steps(n); b.synchronize()
best = None
for _ in range(3):
    b.timer_start(); b.do_steps(n); ms = b.timer_stop(); best = ms if best is None else min(best, ms)
print("size %d  replicas/CTA %d  %.3f ms  %.3e steps/s" % ($S, info["replicas_per_cta"], best, R * n / (best * 1e-3)))
PY
done
