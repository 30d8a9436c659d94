#!/usr/bin/env python
"""Throughput of one model / lattice against the number of replicas in the batch (planner's kernel choice).

    python tools/replica_scaling.py [model] [LxL] [steps] [R ...]
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import quick_perf  # noqa: E402
from kmos_b200 import capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "pairwise_lat_int"
size = [int(x) for x in sys.argv[2].split("x")] if len(sys.argv) > 2 else [128, 128]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
for R in [int(x) for x in sys.argv[4:]] or [1024, 2048, 3552, 7104]:
    out = quick_perf.probe(name, size, R, steps, capi.KERNEL_AUTO)
    print("%s %s R=%d kernel=%s: %.3e kMC steps/s (%.2f ms per %d steps, %d replicas ok)" %
          (name, size, R, out["kernel"], out["steps_per_s"], out["ms"], steps, out["ok_replicas"]), flush=True)
