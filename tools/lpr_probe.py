#!/usr/bin/env python
"""Generated kernel at a given lane-group width: parity against the oracle on a small batch, then throughput.

    python tools/lpr_probe.py <model> <LxL> <R> <steps> <lpr> [<lpr> ...]
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import load_model  # noqa: E402
from util import compare_batch, make_inputs, run_oracles  # noqa: E402
from kmos_b200 import capi, engine, workloads  # noqa: E402

name, size, R, n = sys.argv[1], [int(x) for x in sys.argv[2].split("x")], int(sys.argv[3]), int(sys.argv[4])
ir, blob, info = load_model(name)
for lpr in [int(x) for x in sys.argv[5:]]:
    Rp = 13
    rates, lut, seeds = make_inputs(ir, info, Rp, seed=5)
    b = engine.Batch(engine.Model(ir=ir, blob=blob, info=info), Rp, size, seeds=seeds, rates=rates,
                     kernel=capi.KERNEL_GENERATED, proclist="build", lpr=lpr)
    gen = run_oracles(blob, size, rates, lut, seeds, [1500, 1500])
    compare_batch(b, next(gen), avail_replicas=(0,))
    for k, oracles in zip([1500, 1500], gen):
        b.do_steps(k)
        compare_batch(b, oracles, avail_replicas=(0, Rp - 1))
    b.close()
    m = engine.Model(ir=ir, blob=blob, info=info)
    b = engine.Batch(m, R, size, rates=workloads.rates_for(name, ir, R), kernel=capi.KERNEL_GENERATED,
                     proclist="build", lpr=lpr)
    ki = b.kernel_info()
    b.do_steps(n)
    b.synchronize()
    ts = []
    for _ in range(3):
        b.timer_start()
        b.do_steps(n)
        ts.append(b.timer_stop())
    print("%s %s lpr=%d parity ok; R=%d: %.3e kMC steps/s (%.2f ms per %d steps; %d replicas per CTA, %d regs)" %
          (name, size, lpr, R, R * n / (min(ts) * 1e-3), min(ts), n, ki["replicas_per_cta"], ki["registers"]), flush=True)
    b.close()
