"""--search stage (search_filter) and the sequence comparator (cseq_comparator): the closed-form restatement and the
CUDA kernels against the reference's own cseq_comparator.cpp compiled in place (oracle/_ref)."""
import numpy as np
import pytest

import sina_b200
from oracle import oracle as O
from sina_b200 import synth

RULES = [(iu, co, flc) for iu in (0, 1, 2) for co in range(9) for flc in (False, True)]


def random_aligned(rng, W, n, lowercase=0.0, lo=0, hi=None):
    hi = W if hi is None else hi
    cols = np.sort(rng.choice(np.arange(lo, hi), size=n, replace=False)).astype(np.uint32)
    masks = rng.choice(np.array([1, 2, 4, 8, 1, 2, 4, 8, 1, 2, 4, 8, 3, 5, 15], np.uint8), size=n)
    masks = masks | (16 * (rng.random(n) < lowercase)).astype(np.uint8)
    return masks.astype(np.uint8), cols


def pair_cases(rng, n_cases):
    """pairs with every geometry traverse() distinguishes: overlapping, nested, disjoint either way, touching, equal"""
    W = 400
    for i in range(n_cases):
        kind = i % 6
        lc = [0.0, 0.3, 0.9][i % 3]
        if kind == 0:
            a, b = random_aligned(rng, W, 60, lc), random_aligned(rng, W, 80, lc)
        elif kind == 1:
            a, b = random_aligned(rng, W, 40, lc, 100, 200), random_aligned(rng, W, 120, lc)
        elif kind == 2:
            a, b = random_aligned(rng, W, 30, lc, 0, 150), random_aligned(rng, W, 30, lc, 200, 400)
        elif kind == 3:
            a, b = random_aligned(rng, W, 30, lc, 250, 400), random_aligned(rng, W, 30, lc, 0, 200)
        elif kind == 4:
            a = random_aligned(rng, W, 50, lc)
            b = (rng.permutation(a[0]).astype(np.uint8), a[1].copy())
        else:
            a, b = random_aligned(rng, W, 25, lc, 0, 200), random_aligned(rng, W, 25, lc, 199, 400)
        yield a, b


@pytest.fixture(scope="module")
def ref():
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    return O.Ref()


def test_comparator_closed_form_vs_reference(ref):
    rng = np.random.default_rng(7)
    n = 0
    for (am, ac), (bm, bc) in pair_cases(rng, 90):
        for iu, co, flc in RULES:
            for dist in (0, 1):
                if co == 0 and dist == 1:
                    continue
                if flc and (not ((am & 16) == 0).any() or not ((bm & 16) == 0).any()):
                    continue   # everything filtered: the reference reads past the end
                want = np.float32(ref.compare(am, ac, bm, bc, iu, dist, co, flc))
                got = O.compare_np(am, ac, bm, bc, iu, dist, co, flc)
                assert (np.isnan(want) and np.isnan(got)) or want.view(np.uint32) == got.view(np.uint32), (n, iu, co, flc, dist)
                n += 1
    assert n > 5000


def small_db(rng, N=300, W=1200, L=300):
    tree, m, c, o = synth.synth_msa(N, W=W, L=L, seed=int(rng.integers(1 << 30)))
    names = ["seq%05d" % int(x) for x in rng.permutation(N)]
    return tree, O.MSA(m, c, o, W, names=names)


@pytest.mark.gpu
def test_identity_kernel_vs_reference(ref):
    rng = np.random.default_rng(11)
    W = 400
    rows = [random_aligned(rng, W, int(rng.integers(20, 150)), [0.0, 0.2][i % 2]) for i in range(24)]
    off = np.zeros(len(rows) + 1, np.uint64)
    off[1:] = np.cumsum([len(r[0]) for r in rows])
    msa = O.MSA(np.concatenate([r[0] for r in rows]), np.concatenate([r[1] for r in rows]), off, W)
    ix = sina_b200.Index(msa.masks, msa.cols, msa.off, W, k=4)
    qs = [a for a, _ in pair_cases(rng, 18)]
    aoff = np.zeros(len(qs) + 1, np.uint64)
    aoff[1:] = np.cumsum([len(q[0]) for q in qs])
    am, ac = np.concatenate([q[0] for q in qs]), np.concatenate([q[1] for q in qs])
    ref_ids = np.tile(np.arange(msa.N, dtype=np.uint32), len(qs))
    ref_off = (np.arange(len(qs) + 1) * msa.N).astype(np.uint64)
    for iu, co, flc in RULES:
        for dist in (0, 1):
            if co == 0 and dist == 1:
                continue
            got = ix.identity(am, ac, aoff, ref_ids, ref_off, iu, dist, co, int(flc))
            for qi, (qm, qc) in enumerate(qs):
                for r in range(msa.N):
                    bm, bc = msa.row(r)
                    if flc and (not ((qm & 16) == 0).any() or not ((bm & 16) == 0).any()):
                        continue
                    want = np.float32(ref.compare(qm, qc, bm, bc, iu, dist, co, flc))
                    g = got[qi * msa.N + r]
                    assert (np.isnan(want) and np.isnan(g)) or want.view(np.uint32) == g.view(np.uint32), (qi, r, iu, co, flc, dist)
    ix.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["default", "ignore_super", "pessimistic_all", "few_candidates"])
def test_search_stage_vs_reference(ref, variant):
    """search_filter::operator() on aligned sequences produced by the aligner itself"""
    rng = np.random.default_rng(5)
    tree, msa = small_db(rng)
    qm, qo = synth.synth_queries(tree, 40, "full", seed=3)
    ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=8)
    ix.set_name_ranks(msa.names)
    oc, om, res = ix.run(qm, qo, sina_b200.FamParams(fs_min=10, fs_max=10, fs_min_len=50, fs_full_len=250, fs_req_gaps=0),
                         sina_b200.AlignParams())
    assert (res["status"] == 0).all()
    amasks, acols, aoff = [], [], [0]
    for q in range(len(qo) - 1):
        a, n = int(qo[q]), int(res["n_out"][q])
        amasks.append(om[a:a + n]); acols.append(oc[a:a + n]); aoff.append(aoff[-1] + n)
    # a few references themselves as queries: identity 1, and supersequences for --search-ignore-super
    for r in (3, 77, 150):
        m_, c_ = msa.row(r)
        cut = slice(20, len(m_) - 30) if variant == "ignore_super" else slice(0, len(m_))
        amasks.append(m_[cut]); acols.append(c_[cut]); aoff.append(aoff[-1] + len(m_[cut]))
    # one sequence below the 20-base limit
    amasks.append(msa.row(5)[0][:12]); acols.append(msa.row(5)[1][:12]); aoff.append(aoff[-1] + 12)
    amasks, acols, aoff = np.concatenate(amasks), np.concatenate(acols), np.array(aoff, np.uint64)
    kw = dict(default={}, ignore_super=dict(ignore_super=1, min_sim=0.1),
              pessimistic_all=dict(iupac=1, cover=4, min_sim=0.3, max_result=25),
              few_candidates=dict(kmer_candidates=7, max_result=10, min_sim=0.0))[variant]
    sp = sina_b200.SearchParams(**kw)
    ids, sc, n = ix.search(amasks, acols, aoff, sp)
    db = ref.db(msa)
    rix = ref.kidx_build(db, 8)
    nonempty = 0
    for q in range(len(aoff) - 1):
        a, b = int(aoff[q]), int(aoff[q + 1])
        rid, rsc = ref.search(rix, amasks[a:b], acols[a:b], sp.kmer_candidates, sp.max_result, sp.min_sim, sp.ignore_super,
                              sp.iupac, sp.correction, sp.cover, sp.filter_lowercase)
        assert n[q] == len(rid), (variant, q, n[q], len(rid))
        assert (ids[q, :n[q]] == rid).all(), (variant, q)
        assert (sc[q, :n[q]].view(np.uint32) == rsc.view(np.uint32)).all(), (variant, q)
        nonempty += len(rid) > 0
    assert n[-1] == 0 and nonempty >= 3
    ref.kidx_free(rix)
    ref.db_free(db)
    ix.close()


@pytest.mark.gpu
def test_search_jukes_cantor_unrelated_rows(ref):
    """--search-correction jc. The reference applies jukes_cantor() to the IDENTITY (src/cseq_comparator.cpp:279-287), so
    any candidate more than 75 % identical becomes NaN and its partial_sort is then undefined (NaN breaks the ordering of
    search::result_item); sina_b200 drops such pairs. The comparison is therefore made on unrelated rows, where no NaN
    arises and the reference's order is defined."""
    rng = np.random.default_rng(23)
    W = 600
    rows = [random_aligned(rng, W, int(rng.integers(150, 300))) for _ in range(64)]
    off = np.zeros(len(rows) + 1, np.uint64)
    off[1:] = np.cumsum([len(r[0]) for r in rows])
    names = ["r%03d" % int(x) for x in rng.permutation(len(rows))]
    msa = O.MSA(np.concatenate([r[0] for r in rows]), np.concatenate([r[1] for r in rows]), off, W, names=names)
    ix = sina_b200.Index(msa.masks, msa.cols, msa.off, W, k=4, nofast=True)
    ix.set_name_ranks(names)
    qs = [random_aligned(rng, W, int(rng.integers(100, 300))) for _ in range(12)]
    aoff = np.zeros(len(qs) + 1, np.uint64)
    aoff[1:] = np.cumsum([len(q[0]) for q in qs])
    am, ac = np.concatenate([q[0] for q in qs]), np.concatenate([q[1] for q in qs])
    db = ref.db(msa)
    rix = ref.kidx_build(db, 4, True)
    for cover in (1, 3, 4, 7):
        sp = sina_b200.SearchParams(kmer_candidates=40, max_result=15, min_sim=0.02, correction=1, cover=cover)
        ids, sc, n = ix.search(am, ac, aoff, sp)
        for q, (qm, qc) in enumerate(qs):
            rid, rsc = ref.search(rix, qm, qc, 40, 15, 0.02, 0, 0, 1, cover, 0)
            assert n[q] == len(rid) and len(rid) > 0, (cover, q)
            assert (ids[q, :n[q]] == rid).all(), (cover, q)
            assert (sc[q, :n[q]].view(np.uint32) == rsc.view(np.uint32)).all(), (cover, q)
    ref.kidx_free(rix)
    ref.db_free(db)
    ix.close()


@pytest.mark.gpu
def test_family_identity_filter_vs_reference(ref):
    """--fs-msc-max < 1 (remove_similar, src/famfinder.cpp:553-556): pre-aligned queries, candidates more identical than the
    threshold are skipped. Queries: references themselves with a few columns changed, and aligned outputs."""
    rng = np.random.default_rng(17)
    tree, msa = small_db(rng, N=400)
    names = msa.names
    ix = sina_b200.Index(msa.masks, msa.cols, msa.off, msa.W, k=8)
    qm, qc, qo, strings = [], [], [0], []
    for r in rng.choice(msa.N, 30, replace=False):
        m_, c_ = msa.row(int(r))
        m_ = m_.copy()
        flip = rng.choice(len(m_), size=len(m_) // 15, replace=False)
        m_[flip] = rng.choice(np.array([1, 2, 4, 8], np.uint8), size=len(flip))
        keep = np.sort(rng.choice(len(m_), size=len(m_) - 10, replace=False))
        qm.append(m_[keep]); qc.append(c_[keep]); qo.append(qo[-1] + len(keep))
        s = np.full(msa.W, ord("-"), np.uint8)
        s[c_[keep]] = O.MASK2RNA[m_[keep] & 31]
        strings.append(s.tobytes().decode())
    qm, qc, qo = np.concatenate(qm), np.concatenate(qc), np.array(qo, np.uint64)
    db = ref.db(msa)
    rix = ref.kidx_build(db, 8)
    differs = 0
    for thr in (0.97, 0.9, 0.8):
        kw = dict(fs_min=8, fs_max=8, fs_min_len=50, fs_full_len=250, fs_req_gaps=0, fs_msc_max=thr)
        ids, sc, n = ix.family(qm, qo, sina_b200.FamParams(**kw), qcols=qc)
        ids0, _, n0 = ix.family(qm, qo, sina_b200.FamParams(**dict(kw, fs_msc_max=2.0)))
        for q in range(len(qo) - 1):
            rn, rid, rsc = ref.family(rix, strings[q], O.FamParams(**kw))
            assert n[q] == rn, (thr, q, n[q], rn)
            assert (ids[q, :rn] == rid).all() and (sc[q, :rn] == rsc).all(), (thr, q)
            differs += n0[q] != n[q] or (ids0[q, :n0[q]] != ids[q, :n[q]]).any()
    assert differs > 20     # the filter does remove candidates
    with pytest.raises(sina_b200.SinaB200Error):
        ix.family(qm, qo, sina_b200.FamParams(fs_msc_max=0.9))    # positions are required
    ref.kidx_free(rix)
    ref.db_free(db)
    ix.close()
