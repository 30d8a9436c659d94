#!/usr/bin/env python
"""Throughput of the five BASELINE.json configurations on one GPU (device-resident state, CUDA events):
the parity-test cases of bench.py's headline, measured the same way for DESIGN.md.  Writes one JSON object."""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from kmos_b200 import engine, otf as otf_mod, tables, workloads  # noqa: E402

CONFIGS = [
    ("A mini_101 fcc_100, local_smart", "mini_101_local_smart", [20, 20], 16384, 20000),
    ("B ZGB 64x64, local_smart", "zgb_local_smart", [64, 64], 4096, 4000),
    ("C RuO2 CO oxidation 20x20, local_smart", "ruo2_local_smart", [20, 20], 16384, 5000),
    ("D pairwise interaction 128x128, lat_int", "pairwise_lat_int", [128, 128], 2048, 4000),
    ("E pairwise interaction 256x256, otf", "pairwise_otf_otf", [256, 256], 3552, 40),
]


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="config letters to run, e.g. DE")
    ap.add_argument("--replicas", type=int, default=0, help="override the replica count")
    ap.add_argument("--out", default="configs_r1.json")
    args = ap.parse_args()
    out = []
    for label, name, size, R, n in CONFIGS:
        if args.only and label[0] not in args.only:
            continue
        R = args.replicas or R
        ir = tables.load_ir(os.path.join(REPO, "tests", "golden", "models", name + ".json"))
        m = engine.Model(ir=ir)
        rates = workloads.rates_for(name.split("_")[0], ir, R)
        lut = None
        if ir["backend"] == "otf":
            lut = np.tile(otf_mod.build_lut(ir, m.info, rates[0]), (R, 1))
        b = engine.Batch(m, R, size, rates=rates, lut=lut)
        b.do_steps(max(n // 4, 1))
        b.synchronize()
        best = None
        for _ in range(3):
            b.timer_start()
            b.do_steps(n)
            ms = b.timer_stop()
            best = ms if best is None else min(best, ms)
        info = b.kernel_info()
        out.append({"config": label, "model": name, "lattice": size, "replicas": R, "steps_per_launch": n,
                    "kernel": info["kernel_name"], "replicas_per_cta": info["replicas_per_cta"],
                    "lists_in_l2": info["lists_in_l2"], "ms": best, "kmc_steps_per_s": R * n / (best * 1e-3),
                    "all_ok": bool((b.status == 0).all())})
        print(json.dumps(out[-1]), flush=True)
        b.close()
    with open(os.path.join(REPO, "gpurun_out", args.out), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    main()
