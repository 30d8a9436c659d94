#!/usr/bin/env python
"""Quick throughput probe (not the bench contract): steps/s of both kernels on a few workloads."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from kmos_b200 import capi, engine, tables, workloads  # noqa: E402


def probe(name, size, R, n, kernel, reps=3):
    ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", name + ".json"))
    m = engine.Model(ir=ir)
    rates = workloads.rates_for(name, ir, R)
    b = engine.Batch(m, R, size, rates=rates, kernel=kernel)
    info = b.kernel_info()
    b.do_steps(n)
    b.synchronize()
    best = None
    for _ in range(reps):
        b.timer_start()
        b.do_steps(n)
        ms = b.timer_stop()
        best = ms if best is None else min(best, ms)
    st = b.status
    out = {"model": name, "size": size, "R": R, "n": n, "kernel": info["kernel_name"], "ms": best,
           "steps_per_s": R * n / (best * 1e-3), "ok_replicas": int((st == 0).sum()), "info": info}
    b.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "ruo2"
    cases = {
        "ruo2": [("ruo2_local_smart", [20, 20], 16384, 2000, capi.KERNEL_SMEM),
                 ("ruo2_local_smart", [20, 20], 16384, 200, capi.KERNEL_GENERIC)],
        "planner": [("ruo2_local_smart", [20, 20], 16384, 2000, capi.KERNEL_SMEM),
                    ("ruo2_local_smart", [20, 20], 16384, 2000, capi.KERNEL_WARP_HBM),
                    ("zgb_local_smart", [64, 64], 4096, 2000, capi.KERNEL_SMEM),
                    ("zgb_local_smart", [64, 64], 4096, 2000, capi.KERNEL_WARP_HBM),
                    ("zgb_local_smart", [32, 32], 8192, 2000, capi.KERNEL_SMEM),
                    ("zgb_local_smart", [32, 32], 8192, 2000, capi.KERNEL_WARP_HBM),
                    ("ab_local_smart", [20, 20], 16384, 4000, capi.KERNEL_SMEM),
                    ("ab_local_smart", [20, 20], 16384, 4000, capi.KERNEL_WARP_HBM),
                    ("pairwise_local_smart", [30, 30], 8192, 2000, capi.KERNEL_SMEM),
                    ("pairwise_local_smart", [30, 30], 8192, 2000, capi.KERNEL_WARP_HBM)],
        "many": [("pairwise84_lat_int", [128, 128], 2048, 2000, capi.KERNEL_WARP_HBM),
                 ("pairwise84_lat_int", [128, 128], 2048, 200, capi.KERNEL_GENERIC),
                 ("pairwise84_local_smart", [64, 64], 4096, 2000, capi.KERNEL_WARP_HBM),
                 ("pdopd_local_smart", [20, 20], 4096, 2000, capi.KERNEL_WARP_HBM),
                 ("pdopd_local_smart", [20, 20], 4096, 200, capi.KERNEL_GENERIC)],
        "all": [("mini_101_local_smart", [20, 20], 16384, 5000, capi.KERNEL_SMEM),
                ("zgb_local_smart", [64, 64], 4096, 1000, capi.KERNEL_SMEM),
                ("ruo2_local_smart", [20, 20], 16384, 2000, capi.KERNEL_SMEM),
                ("ruo2_local_smart", [20, 20], 16384, 200, capi.KERNEL_GENERIC),
                ("pairwise_lat_int", [128, 128], 2048, 2000, capi.KERNEL_WARP_HBM),
                ("pairwise_lat_int", [128, 128], 2048, 200, capi.KERNEL_GENERIC),
                ("ruo2_lat_int", [20, 20], 16384, 1000, capi.KERNEL_WARP_HBM),
                ("pairwise_otf_otf", [64, 64], 512, 100, capi.KERNEL_GENERIC)],
    }[which]
    for c in cases:
        t0 = time.time()
        try:
            print(json.dumps(probe(*c)), flush=True)
        except Exception as e:  # keep going: this is a probe
            print(json.dumps({"model": c[0], "error": str(e)}), flush=True)
        print("  wall %.1fs" % (time.time() - t0), flush=True)
    gb, mhz = engine.measure_smem_bandwidth(0)
    print(json.dumps({"smem_bandwidth_GBps": gb, "sm_clock_mhz": mhz}))
