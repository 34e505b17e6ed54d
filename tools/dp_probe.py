#!/usr/bin/env python
"""Quick DP/align-stage probe on one GPU: N refs (small, fast to synthesise), Q full-length or V4 queries,
prints per-stage device times and GCUPS. Used under ncu for kernel captures (keeps the command short).
    python tools/dp_probe.py [--refs 5000] [--queries 1024] [--kind full] [--reps 3]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import sina_b200
from sina_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--refs", type=int, default=5000)
ap.add_argument("--queries", type=int, default=1024)
ap.add_argument("--kind", default="full")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fs-max", type=int, default=40, help="family size (1: a chain graph, every node has one predecessor)")
a = ap.parse_args()
tree, m, c, o = synth.synth_msa(a.refs, W=50000, L=1500, seed=20260117)
qm, qo = synth.synth_queries(tree, a.queries, a.kind, seed=1000)
ix = sina_b200.Index(m, c, o, 50000, k=10)
s = sina_b200.Session(ix, a.queries, int(qo[-1]))
s.upload(qm, qo)
fp, al = sina_b200.FamParams(fs_min=a.fs_max, fs_max=a.fs_max, fs_req_full=min(1, a.fs_max - 1) if a.fs_max < 40 else 1), sina_b200.AlignParams(realign=1)
s.family(fp)
s.align(al)
s.sync()
s.stats(reset=True)
t0 = time.perf_counter()
for _ in range(a.reps):
    s.align(al)
s.sync()
dt = time.perf_counter() - t0
st = s.stats()
print("align wall %.2f ms/rep; stages ms/rep: graph %.2f dp %.2f backtrack %.2f; cells/query %.0f; DP %.1f GCUPS; whole align %.1f GCUPS"
      % (dt / a.reps * 1e3, st["ms_graph"] / a.reps, st["ms_dp"] / a.reps, st["ms_backtrack"] / a.reps,
         st["cells"] / a.reps / a.queries, st["cells"] / (st["ms_dp"] * 1e-3) / 1e9, st["cells"] / dt / 1e9))
