#!/usr/bin/env python
"""Dump the family graphs (node columns, predecessor lists) of a few synthetic queries to an .npz, for offline
analysis of the DP plan (ring columns / bank groups): python tools/dump_graphs.py out.npz [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import sina_b200
from sina_b200 import synth

out, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4
tree, m, c, o = synth.synth_msa(20000, W=50000, L=1500, seed=20260117)
qm, qo = synth.synth_queries(tree, 64, "full", seed=1000)
ix = sina_b200.Index(m, c, o, 50000, k=10)
s = sina_b200.Session(ix, 64, int(qo[-1]))
s.upload(qm, qo)
s.family(sina_b200.FamParams())
s.align(sina_b200.AlignParams())
s.sync()
d = {}
for q in range(n):
    g = s.dump_graph(q)
    for k in ("col", "pred_off", "preds"):
        d["%s_%d" % (k, q)] = g[k]
np.savez_compressed(out, **d)
print("wrote", out)
