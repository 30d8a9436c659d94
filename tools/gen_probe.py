#!/usr/bin/env python
"""Probe of the exporter-generated kernel: parity against the oracle on small batches, then throughput.

    python tools/gen_probe.py [parity] [perf]
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from kmos_b200 import capi, codegen, engine, tables, workloads  # noqa: E402

CASES = [
    ("mini_101_local_smart", [20, 20], 16, [1, 999, 3000]),
    ("ab_local_smart", [20, 20], 24, [500, 2500, 3000]),
    ("zgb_local_smart", [16, 12], 13, [1000, 4000]),
    ("ruo2_local_smart", [20, 20], 32, [2000, 4000, 6000]),
    ("ruo2_local_smart", [5, 7], 7, [3000, 3000]),
    ("pairwise_local_smart", [10, 9], 9, [2000, 2000]),
    ("hop3d_local_smart", [5, 6, 5], 7, [2000, 2000]),
    ("hop1d_local_smart", [17], 6, [2000, 2000]),
]


def parity():
    from util import compare_batch, make_inputs, run_oracles
    from conftest import load_model
    for name, size, R, chunks in CASES:
        ir, blob, info = load_model(name)
        rates, lut, seeds = make_inputs(ir, info, R, seed=len(name))
        model = engine.Model(ir=ir, blob=blob, info=info)
        try:
            batch = engine.Batch(model, R, size, seeds=seeds, rates=rates, proclist="build",
                                 kernel=capi.KERNEL_GENERATED)
            gen = run_oracles(blob, size, rates, lut, seeds, chunks)
            compare_batch(batch, next(gen), avail_replicas=range(min(R, 3)))
            for n, oracles in zip(chunks, gen):
                batch.do_steps(n)
                compare_batch(batch, oracles, avail_replicas=(0, R - 1))
            print("parity ok", name, size, batch.kernel_info()["kernel_name"], flush=True)
            batch.close()
        except Exception as e:  # a probe: report and go on
            print("parity FAIL", name, size, repr(e)[:400], flush=True)


def perf(cases=None, kernels=("generated", "auto")):
    cases = cases or [("ruo2_local_smart", [20, 20], 16384, 5000), ("mini_101_local_smart", [20, 20], 16384, 5000),
                      ("zgb_local_smart", [64, 64], 4096, 2000), ("ab_local_smart", [20, 20], 16384, 4000),
                      ("pairwise_local_smart", [30, 30], 8192, 2000)]
    for name, size, R, n in cases:
        ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", name + ".json"))
        m = engine.Model(ir=ir)
        rates = workloads.rates_for(name, ir, R)
        for kern in kernels:
            try:
                b = engine.Batch(m, R, size, rates=rates, proclist="build" if kern == "generated" else None,
                                 kernel=capi.KERNEL_GENERATED if kern == "generated" else capi.KERNEL_AUTO)
                info = b.kernel_info()
                b.do_steps(n)
                b.synchronize()
                best = None
                for _ in range(3):
                    b.timer_start()
                    b.do_steps(n)
                    ms = b.timer_stop()
                    best = ms if best is None else min(best, ms)
                print(json.dumps({"model": name, "size": size, "R": R, "n": n, "kernel": info["kernel_name"],
                                  "ms": best, "steps_per_s": R * n / (best * 1e-3),
                                  "ok": int((b.status == 0).sum()), "info": info}), flush=True)
                b.close()
            except Exception as e:
                print(json.dumps({"model": name, "kernel": kern, "error": repr(e)[:300]}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "perf"]
    if "parity" in what:
        parity()
    if "perf" in what:
        perf()
    if "ruo2" in what:
        perf([("ruo2_local_smart", [20, 20], 16384, 5000)])
    if "ruo2gen" in what:
        perf([("ruo2_local_smart", [20, 20], 16384, 5000)], kernels=("generated",))
    if "zgb" in what:   # config B on the generated kernel; KMOS_B200_GEN_LPR=8|16|32 picks the lane-group width
        perf([("zgb_local_smart", [64, 64], 4096, 4000)], kernels=("generated", "auto"))
