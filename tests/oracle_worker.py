"""Worker of tests/util.py:oracle_checkpoints: runs the CPU oracle for a list of replicas, chunk by chunk, and
stores their state after every chunk.  usage: oracle_worker.py in.npz out.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402


def main(fin, fout):
    d = np.load(fin)
    blob, size, chunks = d["blob"], [int(x) for x in d["size"]], [int(x) for x in d["chunks"]]
    lat, ps, tm, st, ok = [], [], [], [], []
    luts = d["lut"] if "lut" in d.files else [None] * len(d["ids"])
    for seed, rid, rt, lt in zip(d["seeds"], d["ids"], d["rates"], luts):
        o = oracle.Oracle(blob, size, seed=int(seed), replica=int(rid), rates=rt, lut=lt)
        rows = ([], [], [], [], [])
        for n in chunks:
            o.do_steps(n)
            for row, v in zip(rows, (o.lattice.astype(np.int8), o.procstat.copy(), o.kmc_time, o.kmc_step, o.status[0])):
                row.append(v)
        for acc, row in zip((lat, ps, tm, st, ok), rows):
            acc.append(np.asarray(row))
    np.savez(fout, lattice=np.asarray(lat), procstat=np.asarray(ps), time=np.asarray(tm), step=np.asarray(st),
             status=np.asarray(ok))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
