"""Pins the C restatement against the reference's own sources compiled in place (oracle/_ref).
Skipped where the prebuilt oracle/_ref/libsina_ref.so is absent."""
import numpy as np
import pytest

from oracle import oracle as O
from sina_b200 import synth


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def params_for(it):
    ap = O.AlignParams(overhang=it % 3, lowercase=[0, 2, 1][(it // 3) % 3], fs_weight=[1.0, 0.0, 2.5][(it // 9) % 3],
                       realign=1)
    if it % 5 == 4:
        ap.match_score, ap.mismatch_score, ap.gap_penalty, ap.gap_ext_penalty = 1.7, -0.9, 4.3, 1.1
    return ap


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_graph_mesh_backtrack_random(orc, ref, seed):
    """graph arrays, all seven mesh cell fields (floats bitwise), output strings, head/tail/score."""
    rng = np.random.default_rng(1000 + seed)
    ncells = 0
    for it in range(80):
        rows, q = synth.random_case(rng, lowercase=0.05 if it % 3 == 0 else 0.0)
        msa = O.MSA.from_rows(rows)
        db = ref.db(msa)
        fam = np.arange(msa.N, dtype=np.uint32)
        ap = params_for(it)
        g1, g2 = orc.graph(msa, fam, ap.fs_weight), ref.graph(db, fam, ap.fs_weight)
        assert (g1["V"], g1["E"]) == (g2["V"], g2["E"])
        for k in ("col", "mask", "pred_off", "preds", "first", "last"):
            assert (g1[k] == g2[k]).all(), k
        assert (bits(g1["weight"]) == bits(g2["weight"])).all()
        qm = O.encode(q)
        rr, s2, c2, log, cells = ref.align(db, fam, q, msa.W, ap, want_cells=True)
        r1, c1, m1, famp = orc.align(msa, fam, qm, ap)
        assert rr.status == r1.status
        if rr.status == 0:
            mesh = orc.mesh(msa, famp[:r1.fam_used], (qm & 15) if ap.lowercase != 1 else qm, ap)
            for k in mesh:
                a, b = mesh[k], cells[k]
                if a.dtype == np.float32:
                    a, b = bits(a), bits(b)
                assert (a == b).all(), (it, k)
            ncells += mesh["value"].size
            assert O.render(m1, c1, msa.W) == s2
            assert (c1 == c2).all()
            assert (r1.head, r1.tail, r1.qual) == (rr.head, rr.tail, rr.qual)
            assert bits(r1.score) == bits(rr.score)
        ref.db_free(db)
    assert ncells > 500000


def test_contains_query_paths(orc, ref):
    """aligner pre-steps (src/align.cpp:329-388): copy from identical / longer relative, --realign erase,
    and the libstdc++ partition permutation of the surviving family."""
    rng = np.random.default_rng(5)
    hits = {0: 0, 1: 0, 2: 0}
    for it in range(120):
        rows, q = synth.random_case(rng, F=int(rng.integers(1, 9)), L=int(rng.integers(20, 80)), overhang_p=0.0)
        msa = O.MSA.from_rows(rows)
        # make the query a substring (or the whole) of some rows
        src = int(rng.integers(0, msa.N))
        m, _ = msa.row(src)
        if len(m) < 6:
            continue
        if it % 3 == 0:
            qm = m.copy()
        else:
            a = int(rng.integers(0, len(m) // 2))
            qm = m[a:a + max(4, len(m) // 2)].copy()
        if it % 4 == 0:
            qm = qm | 16  # case-insensitive
        q = O.decode(qm)
        db = ref.db(msa)
        fam = rng.permutation(msa.N).astype(np.uint32)
        ap = O.AlignParams(realign=it % 2)
        rr, s2, c2, log, _ = ref.align(db, fam, q, msa.W, ap)
        r1, c1, m1, famp = orc.align(msa, fam, qm, ap)
        assert rr.status == r1.status, it
        hits[rr.status] = hits.get(rr.status, 0) + 1
        if rr.status in (0, 1):
            assert O.render(m1, c1, msa.W) == s2
            assert (c1 == c2).all()
        if rr.status == 0:
            assert r1.fam_used == rr.fam_used and bits(r1.score) == bits(rr.score)
        ref.db_free(db)
    assert hits[0] > 5 and hits[1] > 5 and hits[2] > 0


def test_kmers_find_family_random(orc, ref):
    rng = np.random.default_rng(9)
    for (N, L, W, k, nofast) in [(150, 120, 300, 4, 0), (400, 250, 600, 6, 1), (400, 250, 600, 7, 0), (250, 300, 800, 10, 0)]:
        tree, m, c, o = synth.synth_msa(N, W=W, L=L, seed=int(rng.integers(1 << 30)))
        msa = O.MSA(m, c, o, W, names=["r%d" % i for i in range(N)])
        db = ref.db(msa)
        rix, oix = ref.kidx_build(db, k, nofast), orc.index_build(msa, k, nofast)
        off, post = orc.index_lists(oix)
        for kmer in rng.integers(0, 4 ** k, 200):
            assert ref.kidx_list_size(rix, int(kmer)) == off[kmer + 1] - off[kmer]
        qm, qo = synth.synth_queries(tree, 10, "full", seed=int(rng.integers(1 << 30)))
        for i in range(10):
            q = qm[int(qo[i]):int(qo[i + 1])]
            qs = O.decode(q)
            for mode in range(4):
                assert (orc.kmers(q, k, mode) == ref.kmers(qs, k, mode)).all()
            for mx in (1, 7, 41, N + 5):
                s1, i1, p1 = orc.find(oix, q, mx)
                s2, i2, p2 = ref.find(rix, qs, mx)
                assert (s1 == s2).all() and (i1 == i2).all() and p1 == p2
            for fp in (O.FamParams(fs_min=5, fs_max=9, fs_min_len=L // 2, fs_full_len=L - 8, fs_req_gaps=3),
                       O.FamParams(fs_min=3, fs_max=20, fs_msc=30.0, fs_min_len=L - 10, fs_full_len=L, fs_req_full=2,
                                   fs_req_gaps=0),
                       O.FamParams(fs_min=40, fs_max=40, fs_min_len=10, fs_full_len=L + 50, fs_req_gaps=10, fs_req=2)):
                n1, f1, sc1 = orc.family(oix, msa, q, fp)
                n2, f2, sc2 = ref.family(rix, qs, fp)
                assert n1 == n2 and (f1 == f2).all() and (sc1 == sc2).all()
        # leave-query-out: query named like a reference
        fp = O.FamParams(fs_min=5, fs_max=9, fs_min_len=10, fs_full_len=L - 8, fs_req_gaps=0, leave_query_out=1)
        for rid in (0, N // 2):
            q, _ = msa.row(rid)
            n1, f1, _ = orc.family(oix, msa, q, fp, exclude_id=rid)
            n2, f2, _ = ref.family(rix, O.decode(q), fp, qname="r%d" % rid)
            assert n1 == n2 and (f1 == f2).all() and rid not in f1
        ref.kidx_free(rix)
        orc.index_free(oix)
        ref.db_free(db)


def test_inverted_lists_count_the_same(ref):
    """vlimap::invert + increment + offset == plain counting (src/idset.h:315-337,367-384;
    kmer_search.cpp:264-266,392-408): the CSR index may ignore inversion."""
    rng = np.random.default_rng(3)
    for size in (1, 255, 256, 257, 1000):
        for fill in (0.0, 0.1, 0.5, 1.0):
            ids = np.nonzero(rng.random(size) < fill)[0].astype(np.uint32)
            r0, s0 = ref.vlimap_increment(size, ids, False)
            r1, s1 = ref.vlimap_increment(size, ids, True)
            plain = np.zeros(size, np.int16)
            plain[ids] = 1
            assert r0 == 0 and r1 == 1
            assert (s0 == plain).all() and (s1 + r1 == plain).all()


def test_fix_duplicate_positions_random(orc, ref):
    rng = np.random.default_rng(17)
    nerr = 0
    for it in range(1500):
        n = int(rng.integers(1, 60))
        width = int(rng.integers(max(2, n - 5), n * 3 + 2))
        steps = rng.choice([0, 0, 0, 1, 1, 1, 2, 7], n)
        pos = np.minimum(np.cumsum(steps) + int(rng.integers(0, 3)), width - 1).astype(np.uint32)
        masks = (1 << rng.integers(0, 4, n)).astype(np.uint8)
        chars = O.MASK2RNA[masks].copy()
        po, co = np.zeros(n, np.uint32), np.zeros(n, np.uint8)
        st2 = ref.L.ref_fix_duplicate_positions(n, pos, chars, width, it % 2, po, co)
        st1, p1, m1 = orc.fix_duplicate_positions(pos, masks, width, it % 2)
        assert st1 == st2
        nerr += st2
        if st2 == 0:
            assert (p1 == po).all() and O.decode(m1) == co.tobytes().decode()
    assert nerr > 0


def test_turn_check_oracle_vs_ref(orc, ref):
    """--turn: famfinder::impl::turn_check (src/famfinder.cpp:344-378) restated in C vs the harness running the
    reference's own cseq::reverse / complement; queries in all four orientations, odd and even lengths, 'all' and
    'revcomp' modes, and a query without any k-mer hit (best stays 0)"""
    tree, m, c, o = synth.synth_msa(500, W=2500, L=500, seed=11)
    msa = O.MSA(m, c, o, 2500)
    qm, qo = synth.synth_queries(tree, 24, "full", seed=9)
    comp = lambda a: (((a & 2) << 1) | ((a & 4) >> 1) | ((a & 1) << 3) | ((a & 8) >> 3) | (a & 16)).astype(np.uint8)
    oix = orc.index_build(msa, 8, 0)
    db = ref.db(msa)
    rix = ref.kidx_build(db, 8, 0)
    seen = set()
    for i in range(24):
        q = qm[int(qo[i]):int(qo[i + 1])]
        if i % 5 == 0:
            q = q[:-1]
        var = [q, q[::-1].copy(), comp(q), comp(q[::-1].copy())][i % 4]
        for allf in (True, False):
            a, asc = orc.turn_check(oix, var, allf)
            b, bsc = ref.turn_check(rix, O.decode(var), allf)
            assert a == b and (asc == bsc).all(), (i, allf, asc, bsc)
            if allf:
                seen.add(a)
                assert a == [0, 1, 2, 3][i % 4]      # turning it back is what scores best
    assert seen == {0, 1, 2, 3}
    none = O.encode("ACGU")                          # shorter than k: no k-mer, all scores 0
    assert orc.turn_check(oix, none, True)[0] == 0 and ref.turn_check(rix, "ACGU", True)[0] == 0
    orc.index_free(oix)
    ref.kidx_free(rix)
    ref.db_free(db)


def test_insertion_forbid_random(orc, ref):
    """--insertion forbid: transition_aspace_aware (src/mesh.h:377-438) through the reference's own compute / backtrack
    against the C restatement: all seven mesh cell fields, strings, columns, score bits. The cases are dense
    alignments (few free columns), so the gaps_max budget changes many results with respect to --insertion shift."""
    rng = np.random.default_rng(4242)
    ncells, differ = 0, 0
    for it in range(90):
        rows, q = synth.random_case(rng, wfac=[1.0, 1.15, 1.4, 2.0][it % 4], indel=[0.05, 0.1, 0.2][it % 3])
        msa = O.MSA.from_rows(rows)
        db = ref.db(msa)
        fam = np.arange(msa.N, dtype=np.uint32)
        ap = params_for(it)
        ap.insertion = 1
        qm = O.encode(q)
        rr, s2, c2, log, cells = ref.align(db, fam, q, msa.W, ap, want_cells=True)
        r1, c1, m1, famp = orc.align(msa, fam, qm, ap)
        assert rr.status == r1.status, it
        if rr.status == 0:
            mesh = orc.mesh(msa, famp[:r1.fam_used], (qm & 15) if ap.lowercase != 1 else qm, ap)
            for k in mesh:
                a, b = mesh[k], cells[k]
                if a.dtype == np.float32:
                    a, b = bits(a), bits(b)
                assert (a == b).all(), (it, k)
            ncells += mesh["value"].size
            assert O.render(m1, c1, msa.W) == s2, it
            assert (c1 == c2).all()
            assert (r1.head, r1.tail, r1.qual) == (rr.head, rr.tail, rr.qual)
            assert bits(r1.score) == bits(rr.score)
            ap.insertion = 0
            r0, c0, m0, _ = orc.align(msa, fam, qm, ap)
            differ += int(r0.status != 0 or len(c0) != len(c1) or (c0 != c1).any())
        ref.db_free(db)
    assert ncells > 300000 and differ >= 10, (ncells, differ)


@pytest.mark.parametrize("seed", [1, 2])
def test_weighted_scheme_random(orc, ref, seed):
    """positional weights (--filter): scoring_scheme_weighted compiled from the reference (src/scoring_schemes.h:166-241,
    chosen in src/align.cpp:409-415) against the restatement: every mesh cell field, strings, head/tail, score bits;
    with transition_simple and with transition_aspace_aware (--insertion forbid)."""
    rng = np.random.default_rng(4000 + seed)
    ncells = 0
    try:
        for it in range(50):
            rows, q = synth.random_case(rng, lowercase=0.05 if it % 3 == 0 else 0.0)
            msa = O.MSA.from_rows(rows)
            # weights as alignment_stats makes them: 1 for sparse columns, 0.5 - log(rate) (0.5 .. 20) elsewhere
            w = np.where(rng.random(msa.W) < 0.3, 1.0, 0.5 - np.log(rng.uniform(1e-6, 0.95, msa.W))).astype(np.float32)
            orc.set_column_weights(w)
            ref.set_column_weights(w)
            db = ref.db(msa)
            fam = np.arange(msa.N, dtype=np.uint32)
            ap = params_for(it)
            ap.insertion = 1 if it % 4 == 3 else 0
            qm = O.encode(q)
            rr, s2, c2, log, cells = ref.align(db, fam, q, msa.W, ap, want_cells=True)
            r1, c1, m1, famp = orc.align(msa, fam, qm, ap)
            assert rr.status == r1.status, it
            if rr.status == 0:
                mesh = orc.mesh(msa, famp[:r1.fam_used], (qm & 15) if ap.lowercase != 1 else qm, ap)
                for k in mesh:
                    a, b = mesh[k], cells[k]
                    if a.dtype == np.float32:
                        a, b = bits(a), bits(b)
                    assert (a == b).all(), (it, k)
                ncells += mesh["value"].size
                assert O.render(m1, c1, msa.W) == s2 and (c1 == c2).all()
                assert (r1.head, r1.tail, r1.qual) == (rr.head, rr.tail, rr.qual)
                assert bits(r1.score) == bits(rr.score)
            ref.db_free(db)
    finally:
        orc.set_column_weights(None)
        ref.set_column_weights(None)
    assert ncells > 200000
