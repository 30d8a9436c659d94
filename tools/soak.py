#!/usr/bin/env python
"""One-off soak: every fixture x every kernel that accepts it, odd lattice sizes, long runs, all replicas compared
with the oracle (lattice, procstat, nr_of_sites bit-exact; kmc_time 1e-12) at two checkpoints.  Not part of the
test suite (minutes of oracle time); prints one line per case and a summary, writes gpurun_out/soak_r2.json.
Round 2: the generated kernel at every lane-group width (gen8 / gen16 / gen32) is part of it."""
import glob
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import load_model  # noqa: E402
from util import make_inputs, oracle_checkpoints  # noqa: E402
from kmos_b200 import capi, devtables, engine  # noqa: E402

KERNELS = {"local_smart": ["gen4", "gen8", "gen16", "gen32", "smem", "warp_hbm", "generic"], "lat_int": ["warp_hbm", "generic"],
           "otf": ["warp_hbm", "generic"]}
KIND = {"smem": capi.KERNEL_SMEM, "warp_hbm": capi.KERNEL_WARP_HBM, "generic": capi.KERNEL_GENERIC,
        "gen4": capi.KERNEL_GENERATED, "gen8": capi.KERNEL_GENERATED, "gen16": capi.KERNEL_GENERATED,
        "gen32": capi.KERNEL_GENERATED}


def main():
    rng = np.random.RandomState(2026)
    cores = len(os.sched_getaffinity(0))
    out, bad = [], 0
    names = sorted(os.path.basename(f)[:-5] for f in glob.glob(os.path.join(REPO, "tests", "golden", "models", "*.json")))
    for name in names:
        ir, blob, info = load_model(name)
        dim, backend = ir["model_dimension"], ir["backend"]
        size = [int(rng.randint(7, 14)) for _ in range(dim)] if dim < 3 else [int(rng.randint(5, 8)) for _ in range(3)]
        R = 48
        steps = 2000 if backend == "otf" else 30000
        if ir["spuck"] > 8 or len(ir["procs"]) > 64:
            steps //= 3
        rates, lut, seeds = make_inputs(ir, info, R, seed=int(rng.randint(1, 1000)))
        chunks = [steps // 3, steps - steps // 3]
        t0 = time.time()
        ref = oracle_checkpoints(blob, size, seeds, rates, chunks, cores, lut=lut)
        t_or = time.time() - t0
        model = engine.Model(ir=ir, blob=blob, info=info)
        for k in KERNELS[backend]:
            try:
                kw = {"lpr": int(k[3:])} if k[3:].isdigit() else {"proclist": None}
                b = engine.Batch(model, R, size, seeds=seeds, rates=rates, lut=lut, kernel=KIND[k], **kw)
            except (capi.KmosB200Error, devtables.Unsupported) as e:
                out.append({"model": name, "kernel": k, "size": size, "skipped": str(e)[-80:]})
                print(json.dumps(out[-1]), flush=True)
                continue
            ok = True
            for c, n in enumerate(chunks):
                b.do_steps(n)
                ok = ok and np.array_equal(b.lattice.astype(np.int8), ref[0][:, c])
                ok = ok and np.array_equal(b.procstat, ref[1][:, c])
                ok = ok and np.allclose(b.kmc_time, ref[2][:, c], rtol=1e-12, atol=0)
                ok = ok and np.array_equal(b.kmc_step, ref[3][:, c]) and np.array_equal(b.status, ref[4][:, c])
            b.close()
            bad += 0 if ok else 1
            out.append({"model": name, "kernel": k, "size": size, "replicas": R, "steps": steps, "ok": bool(ok),
                        "oracle_s": round(t_or, 1)})
            print(json.dumps(out[-1]), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "soak_r2.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("cases %d, failed %d" % (len([o for o in out if "ok" in o]), bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
