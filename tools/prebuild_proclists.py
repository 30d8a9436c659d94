#!/usr/bin/env python
"""Compile the generated proclist modules of the local_smart fixtures into kmos_b200/_proclist_cache
(the cache travels to the GPU box with the snapshot).  python tools/prebuild_proclists.py [lpr ...]"""
import glob
import os
import sys
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from kmos_b200 import codegen, devtables, tables  # noqa: E402


def one(args):
    path, lpr = args
    ir = tables.load_ir(path)
    try:
        so = codegen.build(ir, lpr=lpr)
        return "%s lpr=%s -> %s" % (os.path.basename(path), lpr or "auto", os.path.basename(so))
    except devtables.Unsupported as e:
        return "%s: declined (%s)" % (os.path.basename(path), e)


if __name__ == "__main__":
    lprs = [int(x) for x in sys.argv[1:]] or [None]
    jobs = [(p, l) for p in sorted(glob.glob(os.path.join(REPO, "tests", "golden", "models", "*_local_smart.json")))
            for l in lprs]
    with ThreadPoolExecutor(8) as ex:
        for line in ex.map(one, jobs):
            print(line, flush=True)
