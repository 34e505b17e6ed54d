#!/usr/bin/env python
"""K-mer search probe for ncu captures: one find over Q full-length queries against N references.
    python tools/find_probe.py [--refs 50000] [--queries 2048] [--k 10] [--nofast]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sina_b200
from sina_b200 import synth
ap = argparse.ArgumentParser()
ap.add_argument("--refs", type=int, default=50000)
ap.add_argument("--queries", type=int, default=2048)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--nofast", action="store_true")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--kind", default="full")
a = ap.parse_args()
tree, m, c, o = synth.synth_msa(a.refs, W=50000, L=1500, seed=20260117)
qm, qo = synth.synth_queries(tree, a.queries, a.kind, seed=1000)
ix = sina_b200.Index(m, c, o, 50000, k=a.k, nofast=a.nofast)
s = sina_b200.Session(ix, a.queries, int(qo[-1]))
s.upload(qm, qo)
s.find(41); s.sync(); s.stats(reset=True)
for _ in range(a.reps):
    s.find(41)
s.sync()
st = s.stats()
print("find %.3f ms/rep, %.0f postings/query, index %s" % (st["ms_find"] / a.reps, st["postings"] / a.reps / a.queries, ix.info()))
