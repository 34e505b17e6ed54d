#!/usr/bin/env python
"""File-to-file throughput of the `sina` command line (FASTA in -> device path -> 50 000-column FASTA out):
    python tools/cli_bench.py [--refs 5000] [--queries 20000] [--gpus 1] [--dir /dev/shm/sina_cli]
The reference MSA is synthetic (sina_b200/synth.py); output goes to --dir (tmpfs by default, the records are
50 kB each, so a disk would be what is measured)."""
import argparse, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sina_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--refs", type=int, default=5000)
ap.add_argument("--queries", type=int, default=20000)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--dir", default="/dev/shm/sina_cli")
ap.add_argument("--batch-size", type=int, default=0)
ap.add_argument("--keep-inputs", action="store_true")
ap.add_argument("--out", default="", help="output file (default: out.fasta in --dir); /dev/null measures the pipeline without the file system")
a = ap.parse_args()
os.makedirs(a.dir, exist_ok=True)
W = 50000
tree, m, c, o = synth.synth_msa(a.refs, W=W, L=1500, seed=20260117)
qm, qo = synth.synth_queries(tree, a.queries, "full", seed=1000)
lut = np.frombuffer(b".AGRCMSVUWKDYHBN.agrcmsvuwkdyhbn", np.uint8)
t0 = time.time()
with open(os.path.join(a.dir, "ref.fasta"), "wb") as f:
    row = np.empty(W, np.uint8)
    for i in range(a.refs):
        row[:] = ord("-")
        s, e = int(o[i]), int(o[i + 1])
        row[c[s:e]] = lut[m[s:e]]
        f.write(b">ref%d\n" % i); f.write(row.tobytes()); f.write(b"\n")
with open(os.path.join(a.dir, "q.fasta"), "wb") as f:
    for i in range(a.queries):
        f.write(b">q%d\n" % i); f.write(lut[qm[int(qo[i]):int(qo[i + 1])]].tobytes()); f.write(b"\n")
print("wrote inputs in %.1f s" % (time.time() - t0), flush=True)
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sina_b200", "bin", "sina")
t0 = time.time()
outp = a.out or os.path.join(a.dir, "out.fasta")
r = subprocess.run([exe, "-i", os.path.join(a.dir, "q.fasta"), "-o", outp, "--db",
                    os.path.join(a.dir, "ref.fasta"), "--fs-engine", "internal", "--gpus", str(a.gpus)] + (["--batch-size", str(a.batch_size)] if a.batch_size else []), capture_output=True, text=True, env=dict(os.environ, SINA_B200_TIMING="1"))
print("rc", r.returncode, "wall %.1f s (includes loading the reference and building the index)" % (time.time() - t0))
print("\n".join(r.stderr.strip().splitlines()[-5:]))
print("output", outp, "bytes", os.path.getsize(outp) if os.path.exists(outp) else None)
for fn in ("ref.fasta", "q.fasta", "out.fasta"):
    try: os.remove(os.path.join(a.dir, fn))
    except OSError: pass
