#!/usr/bin/env python
"""Per-region view of an ncu capture taken with --import-source on: consecutive SASS instructions with the
same execution count per kMC step are one region; prints instructions per step, share of the stall samples
and the leading stall reasons.   python tools/ncu_regions.py <rep> <kmc steps in the launch> [min_share]"""
import csv
import subprocess
import sys


def main():
    rep, steps = sys.argv[1], float(sys.argv[2])
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                                  stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]

    def I(r, k):
        v = r[idx[k]]
        return int(v) if v else 0

    cats = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = sum(I(r, "# Samples") for r in data) or 1
    tot_i = sum(I(r, "Instructions Executed") for r in data)
    print("instructions per kMC step: %.1f   samples: %d" % (tot_i / steps, tot_s))
    agg = {c: sum(I(r, c) for r in data) for c in cats}
    print("stalls overall: " + " ".join("%s=%.1f%%" % (k[6:], 100.0 * v / tot_s) for k, v in
                                         sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    runs, cur = [], None
    for n, r in enumerate(data):
        x = I(r, "Instructions Executed") / steps
        key = None if x < 0.02 else round(x, 1)
        if cur is None or cur["key"] != key:
            cur = {"key": key, "start": n, "n": 0, "inst": 0, "s": 0, "c": {c: 0 for c in cats}}
            runs.append(cur)
        cur["n"] += 1
        cur["inst"] += I(r, "Instructions Executed")
        cur["s"] += I(r, "# Samples")
        for c in cats:
            cur["c"][c] += I(r, c)
    for u in runs:
        if u["s"] < thr * tot_s:
            continue
        top = sorted(u["c"].items(), key=lambda kv: -kv[1])[:4]
        print("%5d n=%3d x%-5s inst/step %6.1f  samples %5.2f%%  %s" % (
            u["start"], u["n"], u["key"], u["inst"] / steps, 100.0 * u["s"] / tot_s,
            " ".join("%s=%.1f%%" % (k[6:], 100.0 * v / tot_s) for k, v in top)))


if __name__ == "__main__":
    main()
