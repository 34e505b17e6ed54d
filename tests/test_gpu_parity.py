"""GPU parity: the CUDA path through the C-ABI (sina_b200) against the oracle (oracle/sina_oracle.c) and
the golden vectors produced by the reference's own code. Bit-exact for ids, columns, strings; the DP score
(north star tolerance: 1e-5 relative) is asserted bit-exact as well."""
import numpy as np
import pytest

import sina_b200
from conftest import load_golden
from oracle import oracle as O
from sina_b200 import synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def pack_queries(qs):
    qoff = np.zeros(len(qs) + 1, np.uint64)
    qoff[1:] = np.cumsum([len(q) for q in qs])
    return np.concatenate(qs).astype(np.uint8), qoff


def gpu_align_one(msa, fam, qm, ap_kw, k=4):
    ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=k)
    fam = np.asarray(fam, np.uint32)
    oc, om, res = ix.align(qm, np.array([0, len(qm)], np.uint64), fam, np.array([0, len(fam)], np.uint64),
                           sina_b200.AlignParams(**ap_kw))
    ix.close()
    return oc, om, res[0]


def compare_result(r_gpu, oc, om, r_orc, c_orc, m_orc, W, ctx):
    assert r_gpu["status"] == r_orc.status, ctx
    if r_orc.status in (0, 1):
        n = r_orc.n_out
        assert r_gpu["n_out"] == n, ctx
        assert (oc[:n] == c_orc).all(), ctx
        assert (om[:n] == m_orc).all(), ctx
    if r_orc.status == 0:
        assert (r_gpu["head"], r_gpu["tail"], r_gpu["qual"], r_gpu["n_nodes"]) == (r_orc.head, r_orc.tail, r_orc.qual, r_orc.n_nodes), ctx
        assert bits(r_gpu["score"]) == bits(r_orc.score), ctx
        assert bits(r_gpu["raw"]) == bits(r_orc.raw) and bits(r_gpu["sum_weight"]) == bits(r_orc.sum_weight), ctx
        assert (r_gpu["end_m"], r_gpu["end_s"]) == (r_orc.end_m, r_orc.end_s), ctx


def test_align_golden_cases():
    """every golden case (hand KATs + 120 random small families), one index per case"""
    cases = load_golden("align_cases")
    for i, e in enumerate(cases):
        msa = O.MSA.from_rows(e["rows"])
        qm = O.encode(e["query"])
        oc, om, r = gpu_align_one(msa, np.arange(msa.N), qm, e["params"])
        assert r["status"] == e["status"], i
        if e["status"] in (0, 1):
            assert O.render(om[:r["n_out"]], oc[:r["n_out"]], msa.W) == e["aligned"], i
        if e["status"] == 0:
            assert (r["head"], r["tail"], r["qual"]) == (e["head"], e["tail"], e["qual"]), i
            if e["fam_used"] == msa.N:  # the golden graph size is that of the whole family
                assert r["n_nodes"] == e["n_nodes"], i
            assert r["fam_used"] == e["fam_used"], i
            assert int(bits(r["score"])) == e["score_bits"], i


@pytest.mark.parametrize("graph_path", ["shared", "generic"])
def test_graph_matches_oracle(orc, graph_path, monkeypatch):
    """family graph (nodes, weights, predecessor lists) through the shared-memory column table and through the
    global-scratch fallback"""
    monkeypatch.setenv("SG_GRAPH_GENERIC", "1" if graph_path == "generic" else "0")
    rng = np.random.default_rng(42)
    for it in range(25):
        rows, q = synth.random_case(rng, lowercase=0.05 if it % 2 else 0.0)
        msa = O.MSA.from_rows(rows)
        fsw = [1.0, 0.0, 2.5][it % 3]
        ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=4)
        qm = O.encode(q)
        s = sina_b200.Session(ix, 1, len(qm))
        s.upload(qm, np.array([0, len(qm)], np.uint64))
        s.set_family(np.arange(msa.N, dtype=np.uint32), np.array([0, msa.N], np.uint64))
        s.align(sina_b200.AlignParams(fs_weight=fsw, realign=1))
        _, _, res = s.download_align()
        if res[0]["status"] != 2:
            g = s.dump_graph(0)
            go = orc.graph(msa, np.arange(msa.N), fsw)
            assert (g["V"], g["E"]) == (go["V"], go["E"])
            for k in ("col", "mask", "pred_off", "preds"):
                assert (g[k] == go[k]).all(), (it, k)
            assert (bits(g["weight"]) == bits(go["weight"])).all()
        s.close()
        ix.close()


@pytest.fixture(params=["v2", "generic"])
def dp_mode(request, monkeypatch):
    """run the DP through the specialised kernel (mesh_v2) and through the generic fallback (mesh_v1)"""
    monkeypatch.setenv("SG_DP_GENERIC", "1" if request.param == "generic" else "0")
    return request.param


def test_align_batch_random_vs_oracle(orc, dp_mode):
    """one MSA, many queries with different families in one batch (different graph sizes per CTA)"""
    rng = np.random.default_rng(7)
    tree, m, c, o = synth.synth_msa(300, W=900, L=260, seed=5)
    msa = O.MSA(m, c, o, 900)
    qm, qo = synth.synth_queries(tree, 48, "full", seed=9)
    ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=6)
    fams, foff = [], [0]
    for i in range(48):
        F = int(rng.integers(1, 41))
        fams.append(rng.choice(300, F, replace=False).astype(np.uint32))
        foff.append(foff[-1] + F)
    for ap_kw in (dict(), dict(overhang=2, lowercase=2), dict(overhang=1, match_score=1.7, mismatch_score=-0.9,
                                                              gap_penalty=4.3, gap_ext_penalty=1.1, fs_weight=0.5)):
        oc, om, res = ix.align(qm, qo, np.concatenate(fams), np.array(foff, np.uint64), sina_b200.AlignParams(**ap_kw))
        for i in range(48):
            a, b = int(qo[i]), int(qo[i + 1])
            r1, c1, m1, _ = orc.align(msa, fams[i], qm[a:b], O.AlignParams(**ap_kw))
            compare_result(res[i], oc[a:b], om[a:b], r1, c1, m1, msa.W, (ap_kw, i))
    ix.close()


def test_penalties_beyond_the_value_bound(orc):
    """gap penalties so large that a cell's best candidate can exceed the reference's initial cell value 1000000
    (src/mesh.h:469-473), negative penalties, zero penalties: such batches leave the specialised kernel (which never
    materialises that initial value) for the generic one and still equal the oracle bit for bit"""
    rng = np.random.default_rng(19)
    tree, m, c, o = synth.synth_msa(200, W=900, L=260, seed=6)
    msa = O.MSA(m, c, o, 900)
    qm, qo = synth.synth_queries(tree, 24, "full", seed=4)
    ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=6)
    fams, foff = [], [0]
    for i in range(24):
        F = int(rng.integers(2, 30))
        fams.append(rng.choice(200, F, replace=False).astype(np.uint32))
        foff.append(foff[-1] + F)
    for ap_kw in (dict(gap_penalty=3000.0, gap_ext_penalty=2500.0), dict(gap_penalty=5.0, gap_ext_penalty=1200.0),
                  dict(gap_penalty=-1.0, gap_ext_penalty=2.0), dict(gap_penalty=0.0, gap_ext_penalty=0.0)):
        oc, om, res = ix.align(qm, qo, np.concatenate(fams), np.array(foff, np.uint64), sina_b200.AlignParams(**ap_kw))
        for i in range(24):
            a, b = int(qo[i]), int(qo[i + 1])
            r1, c1, m1, _ = orc.align(msa, fams[i], qm[a:b], O.AlignParams(**ap_kw))
            compare_result(res[i], oc[a:b], om[a:b], r1, c1, m1, msa.W, (ap_kw, i))
    ix.close()


def test_wide_indegree_and_far_edges(orc, dp_mode):
    """in-degree > 8 switches the traceback to 16-bit cells; long gaps force predecessor rows through the
    global spill path (column-rank distance > ring depth); Lq > W hits the reference's runtime_error."""
    rng = np.random.default_rng(3)
    L, W = 400, 1000
    core = np.sort(rng.choice(W - 50, L, replace=False))
    root = rng.integers(0, 4, L)
    rows = []
    for j in range(14):
        s = ["-"] * W
        gap = 5 + 4 * j  # every row's deletion ends at the same column: the next node collects 14 predecessors
        for i in range(L):
            if 200 - gap <= i < 200:
                continue
            s[core[i]] = "AGCU"[root[i] if (i % 37 != j) else (root[i] + 1) % 4]
        rows.append("".join(s))
    msa = O.MSA.from_rows(rows)
    qm = O.encode("".join("AGCU"[x] for x in root[20:380]))
    fam = np.arange(14, dtype=np.uint32)
    g = orc.graph(msa, fam)
    assert np.diff(g["pred_off"]).max() > 8
    for ap_kw in (dict(), dict(overhang=2)):
        oc, om, r = gpu_align_one(msa, fam, qm, ap_kw)
        r1, c1, m1, _ = orc.align(msa, fam, qm, O.AlignParams(**ap_kw))
        compare_result(r, oc, om, r1, c1, m1, W, ap_kw)
    msa2 = O.MSA.from_rows(["AGCUAGCUAGGCU"] * 2)  # Lq > W
    oc, om, r = gpu_align_one(msa2, np.arange(2), O.encode("AGCUAGCUAGGCUAGC"), {})
    assert r["status"] == sina_b200.SG_Q_NOSPACE


@pytest.mark.parametrize("layout", [None, (64, 3), (32, 24), (32, 13), (128, 1)])
def test_index_and_find_vs_oracle(orc, layout, monkeypatch):
    """index lists and find() ranks; `layout` = (sub-tile size, warps per search CTA) forces several sub-tiles per
    CTA and several CTA tiles per query on these small references (production: 4096 x <= 12, or one tile of <= 14), and
    reaches the three variants of the search kernel (<= 12, 13..14, > 14 warps)"""
    if layout:
        monkeypatch.setenv("SG_SUBTILE", str(layout[0]))
        monkeypatch.setenv("SG_TILE_WARPS", str(layout[1]))
    for (N, L, W, k, nofast) in [(300, 220, 500, 6, 0), (300, 220, 500, 6, 1), (500, 400, 900, 8, 0), (200, 300, 700, 10, 0)]:
        tree, m, c, o = synth.synth_msa(N, W=W, L=L, seed=11 + N + k)
        msa = O.MSA(m, c, o, W)
        oix = orc.index_build(msa, k, nofast)
        off, post = orc.index_lists(oix)
        ix = sina_b200.Index(m, c, o, W, k=k, nofast=bool(nofast))
        assert ix.info()["n_postings"] == off[-1]
        rng = np.random.default_rng(k)
        kmers = np.concatenate([rng.integers(0, 4 ** k, 300), np.nonzero(np.diff(off.astype(np.int64)) > 0)[0][:300]]).astype(np.uint32)
        sizes = ix.list_sizes(kmers)
        assert (sizes == (off[kmers + 1] - off[kmers])).all()
        for km in kmers[-20:]:
            assert (ix.posting_list(int(km)) == post[int(off[km]):int(off[km + 1])]).all()
        qm, qo = synth.synth_queries(tree, 24, "full", seed=3)
        for mx in (1, 16, 50, N + 7):
            sc, ids, nres = ix.find(qm, qo, mx)
            for i in range(24):
                s1, i1, _ = orc.find(oix, qm[int(qo[i]):int(qo[i + 1])], mx)
                assert nres[i] == len(s1)
                assert (sc[i, :nres[i]] == s1).all() and (ids[i, :nres[i]] == i1).all(), (N, k, mx, i)
        orc.index_free(oix)
        ix.close()


def test_find_golden():
    for case in load_golden("kmer_cases")["find"]:
        tree, m, c, o = synth.synth_msa(case["N"], W=case["W"], L=case["L"], seed=case["seed"])
        ix = sina_b200.Index(m, c, o, case["W"], k=case["k"], nofast=bool(case["nofast"]))
        qm, qo = pack_queries([O.encode(q["query"]) for q in case["queries"]])
        sc, ids, nres = ix.find(qm, qo, 50)
        fids, fsc, fn = ix.family(qm, qo, sina_b200.FamParams(**case["fam_params"]))
        for i, qe in enumerate(case["queries"]):
            assert sc[i, :nres[i]].tolist() == qe["scores"] and ids[i, :nres[i]].tolist() == qe["ids"]
            assert fn[i] == qe["fam_n"]
            assert fids[i, :max(fn[i], 0)].tolist() == qe["fam_ids"]
            assert fsc[i, :max(fn[i], 0)].tolist() == qe["fam_scores"]
        ix.close()


def test_family_retry_window_and_quotas(orc):
    """quota rules + the 10x retry loop (famfinder.cpp:591-608): short references force wider windows"""
    tree, m, c, o = synth.synth_msa(600, W=700, L=300, seed=21)
    msa = O.MSA(m, c, o, 700)
    oix = orc.index_build(msa, 6, 0)
    ix = sina_b200.Index(m, c, o, 700, k=6)
    qm, qo = synth.synth_queries(tree, 16, "full", seed=2)
    lens = np.diff(o.astype(np.int64))
    for fp_kw in (dict(fs_min=5, fs_max=9, fs_min_len=int(np.percentile(lens, 70)), fs_full_len=int(lens.max()) - 2, fs_req_gaps=3),
                  dict(fs_min=3, fs_max=20, fs_msc=30.0, fs_min_len=10, fs_full_len=int(lens.max()), fs_req_full=3, fs_req_gaps=0),
                  dict(fs_min=40, fs_max=40, fs_min_len=10, fs_full_len=10 ** 6, fs_req_gaps=10, fs_req=2),
                  dict(fs_min=2, fs_max=4, fs_min_len=10 ** 6, fs_full_len=10, fs_req_gaps=0)):
        fids, fsc, fn = ix.family(qm, qo, sina_b200.FamParams(**fp_kw))
        for i in range(16):
            n1, f1, s1 = orc.family(oix, msa, qm[int(qo[i]):int(qo[i + 1])], O.FamParams(**fp_kw))
            assert fn[i] == n1, (fp_kw, i)
            assert (fids[i, :max(n1, 0)] == f1).all() and (fsc[i, :max(n1, 0)] == s1).all()
    fp_kw = dict(fs_min=5, fs_max=9, fs_min_len=10, fs_full_len=250, fs_req_gaps=0, leave_query_out=1)
    qs = [msa.row(r)[0] for r in (0, 17, 333)]
    qm2, qo2 = pack_queries(qs)
    fids, fsc, fn = ix.family(qm2, qo2, sina_b200.FamParams(**fp_kw), exclude_ids=[0, 17, 333])
    for i, rid in enumerate((0, 17, 333)):
        n1, f1, _ = orc.family(oix, msa, qs[i], O.FamParams(**fp_kw), exclude_id=rid)
        assert fn[i] == n1 and (fids[i, :n1] == f1).all() and rid not in fids[i, :n1]
    orc.index_free(oix)
    ix.close()


def test_find_two_level_merge(orc, monkeypatch):
    """windows whose candidates (max x tiles) exceed what one merge CTA sorts in shared memory: the tiles are merged in
    groups, then the groups' winners (find_merge_plan); rank order must not change"""
    monkeypatch.setenv("SG_SUBTILE", "32")          # 12 x 32 = 384 references per tile -> 16 tiles
    tree, m, c, o = synth.synth_msa(6000, W=700, L=300, seed=35)
    msa = O.MSA(m, c, o, 700)
    oix = orc.index_build(msa, 6, 0)
    ix = sina_b200.Index(m, c, o, 700, k=6)
    assert ix.info()["n_tiles"] >= 16
    qm, qo = synth.synth_queries(tree, 6, "full", seed=5)
    for mx in (1000, 1500, 3000):                   # one level (16 x 1000 keys), 2 groups of 10, 4 groups of 5
        sc, ids, nres = ix.find(qm, qo, mx)
        for i in range(6):
            s1, i1, _ = orc.find(oix, qm[int(qo[i]):int(qo[i + 1])], mx)
            assert nres[i] == len(s1) == mx
            assert (sc[i, :mx] == s1).all() and (ids[i, :mx] == i1).all(), (mx, i)
    orc.index_free(oix)
    ix.close()


def test_family_window_beyond_merge_capacity(orc, monkeypatch):
    """the retry loop reaches windows the shared-memory top-k merge cannot hold (max x tiles > 16384): those queries
    are ranked over the whole index (rank_full_kernel), as the reference's loop ends at max_results >= index size
    (famfinder.cpp:591-608). Mixed batch: some queries meet their quotas in the first window, some never do."""
    monkeypatch.setenv("SG_SUBTILE", "32")          # 12 x 32 = 384 references per tile -> 16 tiles
    tree, m, c, o = synth.synth_msa(6000, W=700, L=300, seed=33)
    msa = O.MSA(m, c, o, 700)
    oix = orc.index_build(msa, 6, 0)
    ix = sina_b200.Index(m, c, o, 700, k=6)
    assert ix.info()["n_tiles"] >= 8
    qm, qo = synth.synth_queries(tree, 12, "full", seed=4)
    lens = np.diff(o.astype(np.int64))
    full = int(np.sort(lens)[-25])                   # only 25 references count as full length
    for fp_kw in (dict(fs_min=40, fs_max=40, fs_min_len=10, fs_full_len=full, fs_req_full=2, fs_req_gaps=0),
                  dict(fs_min=40, fs_max=40, fs_min_len=10, fs_full_len=10 ** 6, fs_req_gaps=0),       # quota never met
                  dict(fs_min=10, fs_max=30, fs_min_len=int(lens.max()) + 1, fs_full_len=10, fs_req_gaps=0)):  # nothing long enough
        fids, fsc, fn = ix.family(qm, qo, sina_b200.FamParams(**fp_kw))
        for i in range(12):
            n1, f1, s1 = orc.family(oix, msa, qm[int(qo[i]):int(qo[i + 1])], O.FamParams(**fp_kw))
            assert fn[i] == n1, (fp_kw, i, fn[i], n1)
            assert (fids[i, :max(n1, 0)] == f1).all() and (fsc[i, :max(n1, 0)] == s1).all(), (fp_kw, i)
    # the whole pipeline still runs after such a batch (the session is reused by the host-buffer entry points)
    oc, om, res = ix.run(qm, qo, sina_b200.FamParams(fs_min_len=10, fs_full_len=full, fs_req_full=2, fs_req_gaps=0), sina_b200.AlignParams())
    ores, occ, omm, *_ = orc.run_batch(oix, msa, qm, qo, O.FamParams(fs_min_len=10, fs_full_len=full, fs_req_full=2, fs_req_gaps=0),
                                       O.AlignParams(), nthreads=4)
    for i in range(12):
        a, n1 = int(qo[i]), ores[i].n_out
        assert res[i]["status"] == ores[i].status and (oc[a:a + n1] == occ[a:a + n1]).all(), i
    orc.index_free(oix)
    ix.close()


def test_pipeline_golden():
    case = load_golden("pipeline_case")
    tree, m, c, o = synth.synth_msa(case["N"], W=case["W"], L=case["L"], seed=case["seed"])
    ix = sina_b200.Index(m, c, o, case["W"], k=case["k"])
    qm, qo = pack_queries([O.encode(q["query"]) for q in case["queries"]])
    oc, om, res = ix.run(qm, qo, sina_b200.FamParams(**case["fam_params"]), sina_b200.AlignParams())
    for i, qe in enumerate(case["queries"]):
        a, b = int(qo[i]), int(qo[i + 1])
        assert res[i]["status"] == qe["status"]
        assert oc[a:b].tolist() == qe["cols"], i
        assert (res[i]["head"], res[i]["tail"], res[i]["qual"], res[i]["n_nodes"]) == (qe["head"], qe["tail"], qe["qual"], qe["n_nodes"])
        assert int(bits(res[i]["score"])) == qe["score_bits"]
    ix.close()


def test_full_size_queries_vs_oracle(orc, dp_mode):
    """full-length (~1500 nt) and V4 (~250 nt) queries against 40-member families on a 50 000-column MSA:
    multi-group graphs (V ~ 3000), default parameters, whole path through sg_run_batch."""
    tree, m, c, o = synth.synth_msa(3000, W=50000, L=1500, seed=20260117)
    msa = O.MSA(m, c, o, 50000)
    oix = orc.index_build(msa, 10, 0)
    ix = sina_b200.Index(m, c, o, 50000, k=10)
    for kind, nq in (("full", 24), ("v4", 40)):
        qm, qo = synth.synth_queries(tree, nq, kind, seed=13)
        oc, om, res = ix.run(qm, qo)
        ores, occ, omm, cells, posts, nt = orc.run_batch(oix, msa, qm, qo)
        for i in range(nq):
            a, b = int(qo[i]), int(qo[i + 1])
            compare_result(res[i], oc[a:b], om[a:b], ores[i], occ[a:a + ores[i].n_out], omm[a:a + ores[i].n_out], 50000, (kind, i))
        assert all(r["status"] == 0 for r in res)
    orc.index_free(oix)
    ix.close()


def test_chunk_pipeline_and_arena_retry(orc, monkeypatch):
    """the batch is cut into chunks dealt to several workspaces/streams (SG_BATCH, SG_STREAMS); queries that do
    not fit the traceback arena are redone in further passes (SG_TB_ARENA_MB): results must not depend on any of it"""
    tree, m, c, o = synth.synth_msa(300, W=900, L=260, seed=5)
    msa = O.MSA(m, c, o, 900)
    oix = orc.index_build(msa, 6, 0)
    nq = 61
    qm, qo = synth.synth_queries(tree, nq, "full", seed=19)
    fp_kw = dict(fs_min=20, fs_max=20, fs_min_len=100, fs_full_len=240, fs_req_gaps=5)
    ores, occ, omm, cells, posts, nt = orc.run_batch(oix, msa, qm, qo, O.FamParams(**fp_kw), O.AlignParams())
    for batch, streams, tb_mb in ((7, 3, None), (16, 2, 1), (5, 4, 1), (64, 1, None)):
        monkeypatch.setenv("SG_BATCH", str(batch))
        monkeypatch.setenv("SG_STREAMS", str(streams))
        if tb_mb is None:
            monkeypatch.delenv("SG_TB_ARENA_MB", raising=False)
        else:
            monkeypatch.setenv("SG_TB_ARENA_MB", str(tb_mb))
        ix = sina_b200.Index(m, c, o, 900, k=6)
        oc, om, res = ix.run(qm, qo, sina_b200.FamParams(**fp_kw), sina_b200.AlignParams())
        for i in range(nq):
            a, b = int(qo[i]), int(qo[i + 1])
            compare_result(res[i], oc[a:b], om[a:b], ores[i], occ[a:a + ores[i].n_out], omm[a:a + ores[i].n_out], 900,
                           (batch, streams, tb_mb, i))
        ix.close()
    orc.index_free(oix)


def test_weighted_scoring_vs_oracle(orc):
    """positional column weights (--filter): scoring_scheme_weighted (src/scoring_schemes.h:166-241) through
    sg_index_set_column_weights against the oracle (pinned to the compiled reference's weighted scheme in
    tests/test_oracle_vs_ref.py): small random families incl. --insertion forbid, and a batch of full graphs"""
    rng = np.random.default_rng(77)
    try:
        for it in range(30):
            rows, q = synth.random_case(rng, lowercase=0.05 if it % 3 == 0 else 0.0)
            msa = O.MSA.from_rows(rows)
            w = np.where(rng.random(msa.W) < 0.3, 1.0, 0.5 - np.log(rng.uniform(1e-6, 0.95, msa.W))).astype(np.float32)
            ap_kw = dict(overhang=it % 3, lowercase=[0, 2, 1][(it // 3) % 3], fs_weight=[1.0, 0.0, 2.5][it % 3], realign=1,
                         insertion=1 if it % 4 == 3 else 0)
            if it % 5 == 4:
                ap_kw.update(match_score=1.7, mismatch_score=-0.9, gap_penalty=4.3, gap_ext_penalty=1.1)
            qm = O.encode(q)
            orc.set_column_weights(w)
            r1, c1, m1, _ = orc.align(msa, np.arange(msa.N), qm, O.AlignParams(**ap_kw))
            ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=4)
            ix.set_column_weights(w)
            fam = np.arange(msa.N, dtype=np.uint32)
            oc, om, res = ix.align(qm, np.array([0, len(qm)], np.uint64), fam, np.array([0, msa.N], np.uint64), sina_b200.AlignParams(**ap_kw))
            compare_result(res[0], oc, om, r1, c1, m1, msa.W, (it, ap_kw))
            ix.close()
        # whole pipeline on larger graphs; switching the weights off again restores the simple scheme
        tree, m, c, o = synth.synth_msa(300, W=900, L=260, seed=5)
        msa = O.MSA(m, c, o, 900)
        oix = orc.index_build(msa, 6, 0)
        qm, qo = synth.synth_queries(tree, 24, "full", seed=29)
        fp_kw = dict(fs_min=20, fs_max=20, fs_min_len=100, fs_full_len=240, fs_req_gaps=5)
        w = (0.5 - np.log(rng.uniform(1e-4, 0.95, 900))).astype(np.float32)
        ix = sina_b200.Index(m, c, o, 900, k=6)
        for weights in (w, None):
            orc.set_column_weights(weights)
            ix.set_column_weights(weights)
            oc, om, res = ix.run(qm, qo, sina_b200.FamParams(**fp_kw), sina_b200.AlignParams())
            ores, occ, omm, *_ = orc.run_batch(oix, msa, qm, qo, O.FamParams(**fp_kw), O.AlignParams(), nthreads=4)
            for i in range(24):
                a, b = int(qo[i]), int(qo[i + 1])
                compare_result(res[i], oc[a:b], om[a:b], ores[i], occ[a:a + ores[i].n_out], omm[a:a + ores[i].n_out], 900, (weights is None, i))
        ix.close()
        orc.index_free(oix)
    finally:
        orc.set_column_weights(None)


def test_oversized_query_fails_alone(orc, monkeypatch):
    """per-query soft failure: a query whose traceback does not fit the arena even alone gets status SG_Q_LIMIT; the
    rest of the batch is aligned as usual, and the session stays usable for the next call"""
    tree, m, c, o = synth.synth_msa(300, W=900, L=260, seed=5)
    msa = O.MSA(m, c, o, 900)
    oix = orc.index_build(msa, 6, 0)
    qm, qo = synth.synth_queries(tree, 9, "full", seed=23)
    rng = np.random.default_rng(3)
    big = (1 << rng.integers(0, 4, 6000)).astype(np.uint8)          # 6000 random bases: ~25x the cells of the others
    qs = [qm[int(qo[i]):int(qo[i + 1])] for i in range(9)]
    qs.insert(4, big)
    qm2, qo2 = pack_queries(qs)
    fp_kw = dict(fs_min=20, fs_max=20, fs_min_len=100, fs_full_len=240, fs_req_gaps=5)
    monkeypatch.setenv("SG_TB_ARENA_MB", "1")
    monkeypatch.setenv("SG_BATCH", "4")
    ix = sina_b200.Index(m, c, o, 900, k=6)
    for rep in range(2):   # the second call reuses the cached session
        oc, om, res = ix.run(qm2, qo2, sina_b200.FamParams(**fp_kw), sina_b200.AlignParams())
        assert res[4]["status"] == sina_b200.SG_Q_LIMIT
        ores, occ, omm, *_ = orc.run_batch(oix, msa, qm2, qo2, O.FamParams(**fp_kw), O.AlignParams(), nthreads=4)
        for i in range(10):
            if i == 4:
                continue
            a, b = int(qo2[i]), int(qo2[i + 1])
            compare_result(res[i], oc[a:b], om[a:b], ores[i], occ[a:a + ores[i].n_out], omm[a:a + ores[i].n_out], 900, (rep, i))
    ix.close()
    orc.index_free(oix)


def test_turn_check_vs_oracle(orc):
    """--turn (sg_turn_batch / sg_session_turn, famfinder::turn_check src/famfinder.cpp:344-378): orientation per
    query against the oracle in both modes; the session leaves the batch in the chosen orientation, so family finding
    and alignment afterwards equal the oracle's on the turned queries"""
    tree, m, c, o = synth.synth_msa(600, W=3000, L=600, seed=5)
    msa = O.MSA(m, c, o, 3000)
    qm, qo = synth.synth_queries(tree, 40, "full", seed=3)
    comp = lambda a: (((a & 2) << 1) | ((a & 4) >> 1) | ((a & 1) << 3) | ((a & 8) >> 3) | (a & 16)).astype(np.uint8)
    turned = lambda q, t: comp(q[::-1].copy() if t & 1 else q) if t & 2 else (q[::-1].copy() if t & 1 else q.copy())
    qs = []
    for i in range(40):
        q = qm[int(qo[i]):int(qo[i + 1])]
        if i % 7 == 0:
            q = q[:-1]                       # odd / even lengths
        qs.append(turned(q, i % 4))
    qs.append(O.encode("ACGUAC"))            # no k-mer at all: stays as it is
    qmask, qoff = pack_queries(qs)
    oix = orc.index_build(msa, 8, 0)
    ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=8)
    fp_kw = dict(fs_min=15, fs_max=15, fs_min_len=100, fs_full_len=560, fs_req_gaps=5)
    for mode in ("all", "revcomp"):
        want = np.array([orc.turn_check(oix, q, mode == "all")[0] for q in qs], np.int32)
        got = ix.turn(qmask, qoff, mode)
        assert (got == want).all(), (mode, got, want)
        if mode == "all":
            assert (want[:40] == np.arange(40) % 4).all() and want[40] == 0
        s = sina_b200.Session(ix, len(qs), len(qmask))
        s.upload(qmask, qoff)
        assert (s.turn(mode) == want).all()
        s.family(sina_b200.FamParams(**fp_kw))
        s.align(sina_b200.AlignParams())
        oc, om, res = s.download_align()
        s.close()
        tq, toff = pack_queries([turned(q, int(t)) for q, t in zip(qs, want)])
        ores, occ, omm, cells, posts, nt = orc.run_batch(oix, msa, tq, toff, O.FamParams(**fp_kw), O.AlignParams())
        for i in range(len(qs)):
            a, n = int(qoff[i]), ores[i].n_out
            assert res[i]["status"] == ores[i].status, (mode, i)
            if ores[i].status in (0, 1):
                assert (oc[a:a + n] == occ[a:a + n]).all() and (om[a:a + n] == omm[a:a + n]).all(), (mode, i)
                assert bits(res[i]["score"]) == bits(ores[i].score), (mode, i)
    orc.index_free(oix)
    ix.close()


def test_insertion_forbid_vs_oracle(orc):
    """--insertion forbid (transition_aspace_aware, src/mesh.h:377-438; generic DP kernel): dense alignments where the
    budget of free columns changes the result, all overhang / lowercase / scoring variants, one batch per setting"""
    rng = np.random.default_rng(99)
    differ = 0
    for it in range(24):
        rows, _ = synth.random_case(rng, F=int(rng.integers(2, 10)), wfac=[1.0, 1.2, 1.5, 2.5][it % 4], indel=[0.05, 0.1, 0.2][it % 3])
        msa = O.MSA.from_rows(rows)
        qs = []
        for j in range(6):
            _, q = synth.random_case(rng, F=2, L=int(msa.off[1] - msa.off[0]) if j % 2 else None)
            qs.append(O.encode(q))
        # queries related to the MSA: mutated copies of its rows
        for j in range(6):
            m, _ = msa.row(int(rng.integers(0, msa.N)))
            q = m.copy()
            flip = rng.random(len(q)) < 0.08
            q[flip] = (1 << rng.integers(0, 4, int(flip.sum()))).astype(np.uint8)
            keep = rng.random(len(q)) >= 0.05
            extra = rng.random(len(q)) < 0.08
            out = []
            for b, k, e in zip(q, keep, extra):
                if k:
                    out.append(b)
                if e:
                    out.append(np.uint8(1 << int(rng.integers(0, 4))))
            if len(out) >= 4:
                qs.append(np.array(out, np.uint8))
        qmask, qoff = pack_queries(qs)
        fam = np.tile(np.arange(msa.N, dtype=np.uint32), len(qs))
        foff = (np.arange(len(qs) + 1) * msa.N).astype(np.uint64)
        ap_kw = dict(insertion=1, overhang=it % 3, lowercase=[0, 2, 1][(it // 3) % 3], fs_weight=[1.0, 0.0, 2.5][(it // 9) % 3],
                     realign=1)
        if it % 5 == 4:
            ap_kw.update(match_score=1.7, mismatch_score=-0.9, gap_penalty=4.3, gap_ext_penalty=1.1)
        ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=4)
        oc, om, res = ix.align(qmask, qoff, fam, foff, sina_b200.AlignParams(**ap_kw))
        ix.close()
        for i, q in enumerate(qs):
            r1, c1, m1, _ = orc.align(msa, np.arange(msa.N, dtype=np.uint32), q, O.AlignParams(**ap_kw))
            a = int(qoff[i])
            compare_result(res[i], oc[a:], om[a:], r1, c1, m1, msa.W, (it, i))
            if r1.status == 0:
                kw0 = dict(ap_kw, insertion=0)
                r0, c0, m0, _ = orc.align(msa, np.arange(msa.N, dtype=np.uint32), q, O.AlignParams(**kw0))
                differ += int(len(c0) != len(c1) or (c0 != c1).any())
    assert differ >= 10, differ


def test_host_buffer_streaming_matches_session(orc, monkeypatch):
    """sg_run_batch streams each finished chunk's output through pinned staging into the caller's buffers: the result
    must equal the session path's bulk download for several chunk / stream settings, for caller-owned buffers reused
    across calls of different sizes (staging growth, cached-session reuse) and for calls without output buffers"""
    tree, m, c, o = synth.synth_msa(500, W=1500, L=350, seed=21)
    qm, qo = synth.synth_queries(tree, 75, "full", seed=8)
    fp = sina_b200.FamParams(fs_min=12, fs_max=12, fs_min_len=100, fs_full_len=330, fs_req_gaps=5)
    ap = sina_b200.AlignParams()
    for batch, streams in ((7, 3), (16, 1), (1000, 4)):
        monkeypatch.setenv("SG_BATCH", str(batch))
        monkeypatch.setenv("SG_STREAMS", str(streams))
        ix = sina_b200.Index(m, c, o, 1500, k=8)
        s = sina_b200.Session(ix, 75, len(qm))
        s.upload(qm, qo)
        s.family(fp)
        s.align(ap)
        want_c, want_m, want_r = s.download_align()
        s.run(fp, ap)                                       # sg_session_run = the same two stages in one call
        run_c, run_m, run_r = s.download_align()
        assert (run_r["status"] == want_r["status"]).all() and (bits(run_r["score"]) == bits(want_r["score"])).all()
        assert (run_c == want_c).all() and (run_m == want_m).all()
        s.close()
        out = (np.full(len(qm), 0xFFFFFFFF, np.uint32), np.full(len(qm), 0xFF, np.uint8), np.zeros(75, sina_b200.RESULT_DTYPE))
        for nq in (75, 20, 61):        # big, small, medium: the cached session and the staging slots are reused
            n = int(qo[nq])
            oc, om, res = ix.run(qm[:n], qo[:nq + 1], fp, ap, out=out)
            assert oc is out[0] and om is out[1]
            for q in range(nq):
                a, k = int(qo[q]), int(want_r[q]["n_out"])
                assert res[q]["status"] == want_r[q]["status"], (batch, streams, nq, q)
                if want_r[q]["status"] in (0, 1):
                    assert (oc[a:a + k] == want_c[a:a + k]).all() and (om[a:a + k] == want_m[a:a + k]).all(), (batch, streams, nq, q)
                    assert bits(res[q]["score"]) == bits(want_r[q]["score"])
        oc2, om2, res2 = ix.run(qm, qo, fp, ap)            # library-allocated outputs
        assert (res2["status"] == want_r["status"]).all()
        ix.close()


def test_edge_case_queries_vs_oracle(orc):
    """corner inputs through the whole path (sg_run_batch) against the oracle: queries of 2..12 bases (shorter than k:
    no k-mer at all), all-ambiguous queries, a query much longer than every reference, a query equal to a reference
    (copy path) and one contained in it, in one batch; then a one-row reference with --fs-req-full 0"""
    tree, m, c, o = synth.synth_msa(120, W=900, L=220, seed=31)
    msa = O.MSA(m, c, o, 900)
    rng = np.random.default_rng(17)
    qs = [np.array([1 << int(x) for x in rng.integers(0, 4, n)], np.uint8) for n in (2, 3, 5, 7, 8, 9, 12)]
    qs.append(np.full(40, 15, np.uint8))                                   # NNNN...
    qs.append(np.array([1 << int(x) for x in rng.integers(0, 4, 700)], np.uint8))   # random, 3x longer than the refs
    r5, _ = msa.row(5)
    qs.append(r5.copy())                                                   # identical to a reference
    qs.append(r5[20:150].copy())                                           # contained in it
    fqm, fqo = synth.synth_queries(tree, 6, "full", seed=77)
    for i in range(6):
        q = fqm[int(fqo[i]):int(fqo[i + 1])].copy()
        if i % 2:
            q[::17] = 15                                                   # sprinkle N
        qs.append(q)
    qmask, qoff = pack_queries(qs)
    for fp_kw in (dict(fs_min=10, fs_max=10, fs_min_len=50, fs_full_len=200, fs_req_gaps=5),
                  dict(fs_min=1, fs_max=1, fs_min_len=10, fs_full_len=200, fs_req_gaps=0, fs_req=1)):
        for ap_kw in (dict(), dict(realign=1, overhang=1)):
            ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=8)
            oc, om, res = ix.run(qmask, qoff, sina_b200.FamParams(**fp_kw), sina_b200.AlignParams(**ap_kw))
            ix.close()
            oix = orc.index_build(msa, 8, 0)
            ores, occ, omm, cells, posts, nt = orc.run_batch(oix, msa, qmask, qoff, O.FamParams(**fp_kw), O.AlignParams(**ap_kw))
            orc.index_free(oix)
            for i in range(len(qs)):
                a, n = int(qoff[i]), ores[i].n_out
                assert res[i]["status"] == ores[i].status, (fp_kw, ap_kw, i, res[i]["status"], ores[i].status)
                if ores[i].status in (0, 1):
                    assert res[i]["n_out"] == n
                    assert (oc[a:a + n] == occ[a:a + n]).all() and (om[a:a + n] == omm[a:a + n]).all(), (fp_kw, ap_kw, i)
                if ores[i].status == 0:
                    assert bits(res[i]["score"]) == bits(ores[i].score), (fp_kw, ap_kw, i)
    # a reference of one row
    one = O.MSA(*[x for x in (msa.masks[:int(msa.off[1])], msa.cols[:int(msa.off[1])], msa.off[:2])], 900)
    fp_kw = dict(fs_min=1, fs_max=5, fs_min_len=10, fs_full_len=100, fs_req_gaps=0, fs_req_full=0)
    ix = sina_b200.Index(one.masks, one.cols, one.off, 900, k=8)
    oc, om, res = ix.run(qmask, qoff, sina_b200.FamParams(**fp_kw), sina_b200.AlignParams(realign=1))
    ix.close()
    oix = orc.index_build(one, 8, 0)
    ores, occ, omm, cells, posts, nt = orc.run_batch(oix, one, qmask, qoff, O.FamParams(**fp_kw), O.AlignParams(realign=1))
    orc.index_free(oix)
    for i in range(len(qs)):
        a, n = int(qoff[i]), ores[i].n_out
        assert res[i]["status"] == ores[i].status, ("one", i)
        if ores[i].status in (0, 1):
            assert (oc[a:a + n] == occ[a:a + n]).all() and (om[a:a + n] == omm[a:a + n]).all(), ("one", i)


def test_production_layout_300k_vs_reference():
    """the configuration the numbers are quoted on: production search layout (sub-tiles of 4096 references, tiles of up
    to 12) over a 300 000-row reference (7 tiles), reference defaults, against the COMPILED REFERENCE (oracle/_ref):
    find ranks, families and aligned columns / score bits of 64 full-length and 64 V4 queries"""
    if not O.have_ref():
        pytest.skip("compiled reference (oracle/_ref) not available")
    N = 300000
    tree, m, c, o = synth.synth_msa(N, W=50000, L=1500, seed=20260117)
    msa = O.MSA(m, c, o, 50000)
    ref = O.Ref()
    db = ref.db(msa)
    rix = ref.kidx_build(db, 10, 0)
    ix = sina_b200.Index(m, c, o, 50000, k=10)
    info = ix.info()
    assert info["n_tiles"] >= 4 and info["tile_size"] % 4096 == 0 and info["tile_size"] >= 8 * 4096
    fp, ap = sina_b200.FamParams(), sina_b200.AlignParams()
    for kind, seed in (("full", 1000), ("v4", 1001)):
        qm, qo = synth.synth_queries(tree, 64, kind, seed=seed)
        queries = [O.decode(qm[int(qo[i]):int(qo[i + 1])]) for i in range(64)]
        # find: rank order (score desc, id desc) of the first window
        sc, ids, nres = ix.find(qm, qo, 41)
        for i in range(0, 64, 8):
            s1, i1, _ = ref.find(rix, queries[i], 41)
            assert (ids[i, :41] == i1).all() and (sc[i, :41] == s1).all(), (kind, i)
        # families
        fids, fsc, fn = ix.family(qm, qo, fp)
        for i in range(0, 64, 4):
            n1, f1, s1 = ref.family(rix, queries[i], O.FamParams())
            assert fn[i] == n1 and (fids[i, :n1] == f1).all() and (fsc[i, :n1] == s1).all(), (kind, i)
        # whole path
        oc, om, res = ix.run(qm, qo, fp, ap)
        rres, roc, rqoff, *_ = ref.run_batch(rix, queries, O.FamParams(), O.AlignParams())
        for i in range(64):
            assert res[i]["status"] == rres[i].status, (kind, i)
            if rres[i].status in (0, 1):
                a, b, n = int(qo[i]), int(rqoff[i]), int(res[i]["n_out"])
                assert (oc[a:a + n] == roc[b:b + n]).all(), (kind, i)
            if rres[i].status == 0:
                assert bits(res[i]["score"]) == bits(rres[i].score), (kind, i)
                assert (res[i]["head"], res[i]["tail"], res[i]["qual"]) == (rres[i].head, rres[i].tail, rres[i].qual), (kind, i)
    ref.kidx_free(rix)
    ref.db_free(db)
    ix.close()
