#!/bin/bash
# lane-group width per model: throughput at 8 / 16 / 32 lanes per replica
for M in "pairwise_local_smart 30 30 8192 2000" "zgb_local_smart 32 32 8192 2000" "ab_local_smart 20 20 16384 4000" "mini_101_local_smart 20 20 16384 5000" "ruo2_local_smart 20 20 16384 5000"; do
set -- $M
for L in 8 16 32; do
python - <<PY
import sys
sys.path.insert(0, ".")
from kmos_b200 import capi, engine, tables, workloads
name, size, R, n = "$1", [$2, $3], $4, $5
ir = tables.load_ir("tests/golden/models/%s.json" % name)
m = engine.Model(ir=ir)
b = engine.Batch(m, R, size, rates=workloads.rates_for(name, ir, R), kernel=capi.KERNEL_GENERATED, lpr=$L)
info = b.kernel_info()
b.do_steps(n); b.synchronize()
best = None
for _ in range(3):
    b.timer_start(); b.do_steps(n); ms = b.timer_stop(); best = ms if best is None else min(best, ms)
print("%s lpr %d replicas/SM %d  %.3e steps/s" % (name, $L, info["replicas_per_cta"] * info["ctas_per_sm"], R * n / (best * 1e-3)))
PY
done; done
