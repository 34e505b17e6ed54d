#!/usr/bin/env python
"""K-mer search sweep on one GPU (BASELINE configs[4]): k in {8,10,12}, fast / --fs-kmer-no-fast,
reference sizes N, Q full-length queries; one JSON line per point with the device time of the search kernels
(CUDA events around find_tile + find_merge on the session's stream), the postings scanned and the achieved HBM GB/s
by §8d's algorithmic-bytes formula 4*P + 2*N + 8*max per query.
    python tools/kmer_sweep.py [--refs 50000,200000,500000] [--k 8,10,12] [--queries 4096] [--reps 3] [--out f.jsonl]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import sina_b200
from sina_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--refs", default="50000,200000,500000")
ap.add_argument("--k", default="8,10,12")
ap.add_argument("--queries", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--max", type=int, default=41)
ap.add_argument("--modes", default="fast,nofast")
ap.add_argument("--out", default="")
a = ap.parse_args()
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = open(a.out, "w") if a.out else None
for N in [int(x) for x in a.refs.split(",")]:
    t0 = time.time()
    tree, m, c, o = synth.synth_msa(N, W=50000, L=1500, seed=20260117)
    qm, qo = synth.synth_queries(tree, a.queries, "full", seed=1000)
    t_syn = time.time() - t0
    for k in [int(x) for x in a.k.split(",")]:
        for mode in a.modes.split(","):
            nofast = mode == "nofast"
            line = {"refs": N, "k": k, "mode": mode, "queries": a.queries, "max": a.max, "synth_s": round(t_syn, 1)}
            try:
                t0 = time.time()
                ix = sina_b200.Index(m, c, o, 50000, k=k, nofast=nofast)
                line["index_build_s"] = round(time.time() - t0, 2)
                info = ix.info() if hasattr(ix, "info") else {}
                line["index"] = {kk: int(v) for kk, v in info.items()} if info else {}
                s = sina_b200.Session(ix, a.queries, int(qo[-1]))
                s.upload(qm, qo)
                s.find(a.max)          # warm-up
                s.sync()
                s.stats(reset=True)
                for _ in range(a.reps):
                    s.find(a.max)
                s.sync()
                st = s.stats()
                ms = st["ms_find"] / a.reps
                P = st["postings"] / a.reps
                nbytes = 4.0 * P + (2.0 * N + 8.0 * a.max) * a.queries
                line.update(ms_find=ms, postings_per_query=P / a.queries, queries_per_s=a.queries / (ms * 1e-3),
                            algorithmic_gb=nbytes / 1e9, achieved_gbs=nbytes / (ms * 1e-3) / 1e9, peak_gbs=peak,
                            frac=nbytes / (ms * 1e-3) / 1e9 / peak)
                s.close()
                ix.close()
            except Exception as e:  # a point that does not fit is reported, not fatal
                line["error"] = str(e)[:200]
            js = json.dumps(line)
            print(js, flush=True)
            if out:
                out.write(js + "\n")
                out.flush()
if out:
    out.close()
