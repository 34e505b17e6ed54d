#!/usr/bin/env python
"""K-mer search A/B: the same find(max) over the same synthetic index with several builds of the library, timed and
compared entry by entry (scores and ids of every rank).
    python tools/find_ab.py --refs 500000 --queries 2048 --libs libsina_b200_old.so libsina_b200.so
The MSA is generated once and handed to one worker process per build through /dev/shm."""
import argparse, os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--refs", type=int, default=50000)
ap.add_argument("--queries", type=int, default=2048)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--max", type=int, default=41)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--libs", nargs="+", default=["libsina_b200.so"])
ap.add_argument("--worker", default=None)
ap.add_argument("--shm", default="/dev/shm/find_ab")
a = ap.parse_args()

if a.worker is None:
    from sina_b200 import synth
    tree, m, c, o = synth.synth_msa(a.refs, W=50000, L=1500, seed=20260117)
    os.makedirs(a.shm, exist_ok=True)
    np.save(a.shm + "/m.npy", m); np.save(a.shm + "/c.npy", c); np.save(a.shm + "/o.npy", o)
    for kind in ("full", "v4"):
        qm, qo = synth.synth_queries(tree, a.queries, kind, seed=1000)
        np.save(a.shm + f"/qm_{kind}.npy", qm); np.save(a.shm + f"/qo_{kind}.npy", qo)
    del m, c, o
    for lib in a.libs:
        env = dict(os.environ, SINA_B200_LIB=lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", lib, "--refs", str(a.refs), "--queries", str(a.queries),
                        "--k", str(a.k), "--max", str(a.max), "--reps", str(a.reps), "--shm", a.shm], env=env, check=True)
    base = a.libs[0]
    for lib in a.libs[1:]:
        for kind in ("full", "v4"):
            r0, r1 = np.load(f"{a.shm}/res_{base}_{kind}.npz"), np.load(f"{a.shm}/res_{lib}_{kind}.npz")
            same = all(np.array_equal(r0[f], r1[f]) for f in ("sc", "ids", "nres"))
            print(f"{kind}: {lib} vs {base}: {'identical' if same else 'DIFFERENT'} ({r0['ids'].size} entries)")
    sys.exit(0)

import sina_b200
m, c, o = (np.load(a.shm + f"/{x}.npy") for x in "mco")
ix = sina_b200.Index(m, c, o, 50000, k=a.k)
for kind in ("full", "v4"):
    qm, qo = np.load(a.shm + f"/qm_{kind}.npy"), np.load(a.shm + f"/qo_{kind}.npy")
    s = sina_b200.Session(ix, a.queries, int(qo[-1]))
    s.upload(qm, qo)
    s.find(a.max); s.sync(); s.stats(reset=True)
    for _ in range(a.reps):
        s.find(a.max)
    s.sync()
    st = s.stats()
    sc, ids, nres = s.download_find()
    np.savez(f"{a.shm}/res_{a.worker}_{kind}.npz", sc=sc, ids=ids, nres=nres)
    print(json.dumps(dict(lib=a.worker, refs=a.refs, kind=kind, queries=a.queries, ms_find=st["ms_find"] / a.reps,
                          us_per_query=1e3 * st["ms_find"] / a.reps / a.queries,
                          postings_per_query=st["postings"] / a.reps / a.queries, tiles=ix.info()["n_tiles"])), flush=True)
    s.close()
