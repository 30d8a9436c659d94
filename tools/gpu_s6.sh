#!/bin/bash
# strong-scaling shard sizes: lane-group width vs replicas per GPU
for R in 2048 4096 8192; do for L in 32 16 8; do
KMOS_B200_GEN_LPR=$L python - <<PY
import os, sys, json
sys.path.insert(0, ".")
from kmos_b200 import capi, engine, tables, workloads
ir = tables.load_ir("tests/golden/models/ruo2_local_smart.json")
m = engine.Model(ir=ir)
R, n = $R, 5000
b = engine.Batch(m, R, [20, 20], rates=workloads.rates_for("ruo2", ir, 16384)[:: 16384 // R][:R].copy(), kernel=capi.KERNEL_GENERATED, lpr=$L)
b.do_steps(n); b.synchronize()
best = None
for _ in range(3):
    b.timer_start(); b.do_steps(n); ms = b.timer_stop(); best = ms if best is None else min(best, ms)
print("R=%d lpr=%d  %.3f ms  %.3e steps/s" % (R, $L, best, R * n / (best * 1e-3)))
PY
done; done
