#!/usr/bin/env python
"""Per-source-line view of an ncu capture taken with --import-source on (code compiled with -lineinfo): share of
the stall samples and instructions per kMC step for the hottest source lines.
    python tools/ncu_lines.py <rep> <kmc steps in the launch> [lines to print]"""
import csv
import subprocess
import sys


def main():
    rep, steps = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                                  stderr=subprocess.DEVNULL).decode()
    cur, hdr, agg = None, None, []
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 4 and r[0] == "Line No":
            hdr = r
            i_s, i_i = r.index("# Samples"), r.index("Instructions Executed")
        elif hdr and len(r) > i_s and r[0].isdigit():
            try:
                agg.append((int(r[i_s] or 0), int(r[i_i] or 0), cur, int(r[0]), r[1].strip()[:96]))
            except ValueError:
                pass
    tot = sum(a[0] for a in agg) or 1
    print("samples: %d   instructions per kMC step: %.1f" % (tot, sum(a[1] for a in agg) / steps))
    for s, i, f, ln, src in sorted(agg, reverse=True)[:top]:
        print("%5.2f%%  inst/step %7.1f  %s:%d  %s" % (100.0 * s / tot, i / steps, f, ln, src))


if __name__ == "__main__":
    main()
