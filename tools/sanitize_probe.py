#!/usr/bin/env python
"""Tiny workloads of every kernel family for `compute-sanitizer --tool memcheck` (one-off check)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import load_model  # noqa: E402
from util import make_inputs  # noqa: E402
from kmos_b200 import capi, engine  # noqa: E402

CASES = [("ruo2_local_smart", [9, 7], capi.KERNEL_GENERATED), ("zgb_local_smart", [12, 11], capi.KERNEL_GENERATED),
         ("mini_101_local_smart", [6, 5], capi.KERNEL_GENERATED), ("pairwise_local_smart", [10, 9], capi.KERNEL_GENERATED),
         ("pairwise_otf_otf", [24, 20], capi.KERNEL_OTF_FAST), ("intzgb_otf", [20, 18], capi.KERNEL_OTF_FAST),
         ("ruo2default_otf", [20, 20], capi.KERNEL_OTF_FAST), ("hop3d_otf", [8, 7, 6], capi.KERNEL_OTF_FAST),
         ("multidentate_otf", [20, 18], capi.KERNEL_OTF_FAST), ("zgb_otf", [24, 22], capi.KERNEL_OTF_FAST),
         ("ab_otf", [20, 20], capi.KERNEL_OTF_FAST),
         ("ruo2_local_smart", [9, 7], capi.KERNEL_SMEM), ("zgb_local_smart", [30, 30], capi.KERNEL_SMEM),
         ("ruo2_local_smart", [9, 7], capi.KERNEL_WARP_HBM), ("pairwise_lat_int", [9, 8], capi.KERNEL_WARP_HBM),
         ("pairwise84_lat_int", [9, 8], capi.KERNEL_WARP_HBM), ("pdopd_local_smart", [6, 5], capi.KERNEL_WARP_HBM),
         ("pairwise_otf_otf", [20, 17], capi.KERNEL_WARP_HBM), ("ruo2default_otf", [8, 7], capi.KERNEL_WARP_HBM),
         ("hop3d_local_smart", [5, 6, 5], capi.KERNEL_SMEM), ("ruo2_lat_int", [6, 6], capi.KERNEL_GENERIC)]
ONLY = sys.argv[1] if len(sys.argv) > 1 else None   # e.g. "otf": cases whose model name contains it
for name, size, kind in CASES:
    if ONLY and ONLY not in name:
        continue
    ir, blob, info = load_model(name)
    R = 11
    rates, lut, seeds = make_inputs(ir, info, R, seed=3)
    lprs = (8, 16, 32) if kind == capi.KERNEL_GENERATED else (None,)
    for lpr in lprs[:-1]:
        g = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, size, seeds=seeds, rates=rates, lut=lut,
                         kernel=kind, lpr=lpr)
        g.do_steps(300)
        g.do_steps(7)
        print(name, "generated lpr", lpr, int((g.status == 0).sum()), int(g.kmc_step.sum()), flush=True)
        g.close()
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), R, size, seeds=seeds, rates=rates, lut=lut, kernel=kind,
                     lpr=lprs[-1])
    b.do_steps(300)
    b.do_steps(7)
    _ = b.avail_sites(0)
    print(name, b.kernel_info()["kernel_name"], int((b.status == 0).sum()), int(b.kmc_step.sum()), flush=True)
    b.close()
