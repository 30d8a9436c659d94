"""The headline workload (RuO2 20x20, 16 384 replicas per GPU) on ONE process driving every GPU through
kmos_b200_fleet_* (engine.Fleet): aggregate kMC steps/s by host wall clock, synchronised on both sides, next to
the same replicas on a single batch.  Not a bench value (bench.py under torchrun is); it shows that one caller
thread keeps the GPUs stepping concurrently.

    python tools/fleet_probe.py [n_gpus] [kmc_steps_per_launch] [launches]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from kmos_b200 import capi, engine  # noqa: E402


def main():
    n_dev = capi.lib().kmos_b200_device_count()
    G = int(sys.argv[1]) if len(sys.argv) > 1 else n_dev
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    launches = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    ir, blob, info, rates, group_of, grid = bench.load_workload()
    model = engine.Model(ir=ir, blob=blob, info=info)
    per = bench.REPLICAS_PER_GPU
    out = {"devices_visible": n_dev, "kmc_steps_per_launch": n, "launches": launches}

    def timed(obj, R):
        for _ in range(3):
            obj.do_steps(n)
        obj.synchronize()
        t0 = time.perf_counter()
        for _ in range(launches):
            obj.do_steps(n)
        obj.synchronize()
        dt = time.perf_counter() - t0
        assert np.all(obj.status == 0)
        return R * n * launches / dt

    for g in sorted({1, G}):
        R = per * g
        gid = np.arange(R)
        seeds = gid.astype(np.uint64) * np.uint64(2654435761) + np.uint64(17)
        rows = np.ascontiguousarray(rates[gid % per])
        fleet = engine.Fleet(model, R, bench.SIZE, gpu_ids=[k % max(n_dev, 1) for k in range(g)], seeds=seeds, rates=rows)
        v = timed(fleet, R)
        t = fleet.split_tally(fleet.reduce_tallies(np.ascontiguousarray(group_of[gid % per]), int(group_of.max()) + 1))
        assert int(t["n_replicas"].sum()) == R and int(t["kmc_steps"].sum()) == R * n * (launches + 3)
        out["fleet_%d_gpu" % g] = {"replicas": R, "kmc_steps_per_s": v, "kernel": fleet.kernel_info()[0]["kernel_name"]}
        fleet.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
