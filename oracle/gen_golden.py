#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference's own code (oracle/_ref/libsina_ref.so).

Run in the build container (needs /root/reference to have built _ref):  python oracle/gen_golden.py
The reference ships no golden alignment for mseq / mesh / backtrack / fix_duplicate_positions
(SURVEY.md §4), so the vectors are produced by running its unmodified sources on seeded inputs.
Every expected value in the JSON files comes from `Ref`; nothing from the restatement or the product.
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from sina_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def params_for(it):
    ap = dict(overhang=it % 3, lowercase=[0, 2, 1][(it // 3) % 3], fs_weight=[1.0, 0.0, 2.5][(it // 9) % 3], realign=1)
    if it % 5 == 4:
        ap.update(match_score=1.7, mismatch_score=-0.9, gap_penalty=4.3, gap_ext_penalty=1.1)
    return ap


def align_cases(ref):
    rng = np.random.default_rng(20260117)
    out = []
    # hand KATs of SURVEY.md Appendix B (#1-#7)
    kats = [
        (["AGCU-AGCUAGG--CU", "AGCUUAGC-AGGA-CU", "-GCU-AGCUCGG--CU"], "AGCUAGCAGGACU", {}),
        (["--AGCUAGCUAGGCU---", "--AGCUAGC-AGGCU---"], "GGAGCUAGCUAGGCUAA", {}),
        (["--AGCUAGCUAGGCU---", "--AGCUAGC-AGGCU---"], "GGAGCUAGCUAGGCUAA", dict(overhang=2, lowercase=2)),
        (["--AGCUAGCUAGGCU---", "--AGCUAGC-AGGCU---"], "GGAGCUAGCUAGGCUAA", dict(overhang=1)),
        (["--AGCUAGCUAGGCU-"] * 2, "AGCUAGCCCCUAGGCU", {}),
        (["AGCU--AGCU--AGGCU"] * 2, "AGCUAGCCCUAGGCU", {}),
        (["AGCUAGCUAGGCUAGCUAGCU"] * 2, "AGCUAGCUAGCUAGCU", {}),
        (["AGCUAGCUAGGCU"] * 2, "AGCUAGCUAGGCUAGC", {}),
        # contains-query shortcut (no --realign): exact copy, substring copy; and --realign dropping all
        (["--AGCUAGCUAGGCU---", "--AGCUAGC-AGGCU---"], "AGCUAGCAGGCU", dict(realign=0)),
        (["--AGCUAGCUAGGCU---", "--AGCUAGC-AGGCU---"], "GCUAGCUAGG", dict(realign=0)),
        (["--AGCUAGCUAGGCU---", "--AGCUaGC-AGGCU---"], "gcuagcuagg", dict(realign=1)),
        (["--AGCUAGCUAGGCU---"], "GCUAGCUAGG", dict(realign=1)),
    ]
    cases = [(rows, q, ap) for rows, q, ap in kats]
    for it in range(120):
        rows, q = synth.random_case(rng, lowercase=0.05 if it % 3 == 0 else 0.0)
        cases.append((rows, q, params_for(it)))
    for rows, q, apd in cases:
        msa = O.MSA.from_rows(rows)
        db = ref.db(msa)
        fam = np.arange(msa.N, dtype=np.uint32)
        ap = O.AlignParams(**apd)
        r, s, cols, log, cells = ref.align(db, fam, q, msa.W, ap, want_cells=True)
        g = ref.graph(db, fam, ap.fs_weight)
        e = dict(rows=rows, query=q, params=apd, status=r.status, aligned=s, head=r.head, tail=r.tail, qual=r.qual,
                 score_bits=int(np.float32(r.score).view(np.uint32)), n_nodes=int(g["V"]), n_edges=int(g["E"]),
                 graph_crc=[crc(g[k]) for k in ("col", "mask", "weight", "pred_off", "preds", "first", "last")])
        if cells is not None:
            e["mesh_crc"] = {k: crc(v) for k, v in cells.items()}
            e["fam_used"] = int(r.fam_used)
        out.append(e)
        ref.db_free(db)
    return out


def fixdup_cases(ref):
    rng = np.random.default_rng(99)
    out = []
    for it in range(200):
        n = int(rng.integers(1, 40))
        width = int(rng.integers(max(2, n - 3), n * 3 + 2))
        # monotone positions with duplicates (insertions), sometimes crowded
        steps = rng.choice([0, 0, 1, 1, 1, 2, 5], n)
        pos = np.minimum(np.cumsum(steps) + int(rng.integers(0, 3)), width - 1).astype(np.uint32)
        pos = np.maximum.accumulate(pos)
        masks = (1 << rng.integers(0, 4, n)).astype(np.uint8)
        chars = O.MASK2RNA[masks].copy()
        po = np.zeros(n, np.uint32)
        co = np.zeros(n, np.uint8)
        lc = it % 2
        st = ref.L.ref_fix_duplicate_positions(n, pos, chars, width, lc, po, co)
        out.append(dict(pos=pos.tolist(), bases=chars.tobytes().decode(), width=width, lowercase=lc, status=st,
                        out_pos=po.tolist() if st == 0 else None,
                        out_bases=co.tobytes().decode() if st == 0 else None))
    return out


def kmer_cases(ref):
    out = dict(kmers=[], find=[])
    rng = np.random.default_rng(5)
    seqs = ["AGCTAGCA", "AGCTNAGCTAGCTNAGCTAGCTAGCTN", "ACGU", "", "A", "AAAAAAAAAAAAAAAAAAAAAA",
            "ACGTRACGTACGTYAAACCCGGGTTTAAA"]
    for _ in range(6):
        seqs.append("".join("AGCU"[x] for x in rng.integers(0, 4, int(rng.integers(30, 200)))))
    for s in seqs:
        for k in (1, 2, 4, 8, 10, 12):
            for mode in range(4):
                out["kmers"].append(dict(seq=s, k=k, mode=mode, kmers=ref.kmers(s, k, mode).tolist()))
    # find / family on a small tree-structured MSA
    for (N, L, W, k, nofast) in [(300, 220, 500, 6, 0), (300, 220, 500, 6, 1), (500, 400, 900, 8, 0),
                                 (200, 300, 700, 10, 0), (64, 150, 400, 4, 0)]:
        tree, m, c, o = synth.synth_msa(N, W=W, L=L, seed=11 + N + k)
        msa = O.MSA(m, c, o, W)
        db = ref.db(msa)
        ix = ref.kidx_build(db, k, nofast)
        qm, qo = synth.synth_queries(tree, 12, "full", seed=3)
        case = dict(N=N, L=L, W=W, k=k, nofast=nofast, seed=11 + N + k, qseed=3, queries=[])
        for i in range(12):
            q = O.decode(qm[int(qo[i]):int(qo[i + 1])])
            sc, ids, P = ref.find(ix, q, 50)
            fp = O.FamParams(fs_min=10, fs_max=15, fs_min_len=L // 2, fs_full_len=L - 12, fs_req_gaps=5)
            nf, fid, fsc = ref.family(ix, q, fp)
            case["queries"].append(dict(query=q, scores=sc.tolist(), ids=ids.tolist(), postings=int(P),
                                        fam_n=int(nf), fam_ids=fid.tolist(), fam_scores=fsc.tolist()))
        case["fam_params"] = dict(fs_min=10, fs_max=15, fs_min_len=L // 2, fs_full_len=L - 12, fs_req_gaps=5)
        out["find"].append(case)
        ref.kidx_free(ix)
        ref.db_free(db)
    return out


def pipeline_case(ref):
    """whole path on a small SILVA-like set: family finding + alignment, reference defaults."""
    N, L, W, k = 400, 300, 1200, 8
    tree, m, c, o = synth.synth_msa(N, W=W, L=L, seed=77)
    msa = O.MSA(m, c, o, W)
    db = ref.db(msa)
    ix = ref.kidx_build(db, k, 0)
    qm, qo = synth.synth_queries(tree, 24, "full", seed=5)
    queries = [O.decode(qm[int(qo[i]):int(qo[i + 1])]) for i in range(24)]
    fpd = dict(fs_min=15, fs_max=15, fs_min_len=100, fs_full_len=280, fs_req_gaps=5)
    fp = O.FamParams(**fpd)
    res, oc, qoff, cells, posts, nt = ref.run_batch(ix, queries, fp, O.AlignParams(), nthreads=2)
    out = dict(N=N, L=L, W=W, k=k, seed=77, qseed=5, nq=24, fam_params=fpd, cells=int(cells), postings=int(posts),
               queries=[])
    for i, q in enumerate(queries):
        a, b = int(qoff[i]), int(qoff[i + 1])
        out["queries"].append(dict(query=q, status=res[i].status, cols=oc[a:b].tolist(), head=res[i].head,
                                   tail=res[i].tail, qual=res[i].qual, n_nodes=int(res[i].n_nodes),
                                   score_bits=int(np.float32(res[i].score).view(np.uint32))))
    ref.kidx_free(ix)
    ref.db_free(db)
    return out


def main():
    ref = O.Ref()
    os.makedirs(GOLD, exist_ok=True)
    for name, fn in [("align_cases", align_cases), ("fixdup_cases", fixdup_cases), ("kmer_cases", kmer_cases),
                     ("pipeline_case", pipeline_case)]:
        data = fn(ref)
        with open(os.path.join(GOLD, name + ".json"), "w") as f:
            json.dump(data, f, separators=(",", ":"))
        print(name, os.path.getsize(os.path.join(GOLD, name + ".json")))


if __name__ == "__main__":
    main()
