"""TEST INFRASTRUCTURE ONLY -- ctypes access to the two checkers.

* ``Oracle``: plain-C restatement of the hot path (oracle/sina_oracle.c -> oracle/liboracle.so)
* ``Ref``:    the reference's own sources compiled in place (oracle/ref_harness.cpp ->
              oracle/_ref/libsina_ref.so; built only where /root/reference exists, shipped prebuilt)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Data layout shared with the product's C-ABI: an MSA is (masks u8[], cols u32[], off u64[N+1], W);
masks are SINA's 4-bit IUPAC masks (A=1 G=2 C=4 U=8, +16 lowercase; src/aligned_base.h:38-52).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsina_ref.so")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C")


def build(ref=True):
    """Compile liboracle.so (always) and _ref/libsina_ref.so (when the reference tree is present)."""
    subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class FamParams(C.Structure):
    """famfinder options (src/famfinder.cpp:155-195), reference defaults."""
    _fields_ = [("fs_min", C.c_uint32), ("fs_max", C.c_uint32), ("fs_msc", C.c_float), ("fs_msc_max", C.c_float),
                ("fs_min_len", C.c_uint32), ("fs_req_full", C.c_uint32), ("fs_full_len", C.c_uint32),
                ("fs_req_gaps", C.c_uint32), ("fs_req", C.c_uint32), ("leave_query_out", C.c_int)]

    def __init__(self, fs_min=40, fs_max=40, fs_msc=0.7, fs_msc_max=2.0, fs_min_len=150, fs_req_full=1,
                 fs_full_len=1400, fs_req_gaps=10, fs_req=1, leave_query_out=0):
        super().__init__(fs_min, fs_max, fs_msc, fs_msc_max, fs_min_len, fs_req_full, fs_full_len, fs_req_gaps,
                         fs_req, leave_query_out)


class AlignParams(C.Structure):
    """aligner options (src/align.cpp:232-259), reference defaults."""
    _fields_ = [("match_score", C.c_float), ("mismatch_score", C.c_float), ("gap_penalty", C.c_float),
                ("gap_ext_penalty", C.c_float), ("fs_weight", C.c_float), ("overhang", C.c_int),
                ("lowercase", C.c_int), ("insertion", C.c_int), ("realign", C.c_int)]

    def __init__(self, match_score=2.0, mismatch_score=-1.0, gap_penalty=5.0, gap_ext_penalty=2.0, fs_weight=1.0,
                 overhang=0, lowercase=0, insertion=0, realign=0):
        super().__init__(match_score, mismatch_score, gap_penalty, gap_ext_penalty, fs_weight, overhang, lowercase,
                         insertion, realign)


class OracleResult(C.Structure):
    _fields_ = [("status", C.c_int), ("score", C.c_float), ("raw", C.c_float), ("sum_weight", C.c_float),
                ("head", C.c_int), ("tail", C.c_int), ("qual", C.c_int), ("n_nodes", C.c_uint32),
                ("fam_used", C.c_uint32), ("n_out", C.c_uint32), ("end_m", C.c_uint32), ("end_s", C.c_uint32)]


class RefResult(C.Structure):
    _fields_ = [("status", C.c_int), ("score", C.c_float), ("head", C.c_int), ("tail", C.c_int), ("qual", C.c_int),
                ("n_nodes", C.c_uint32), ("fam_used", C.c_uint32)]


class _Graph(C.Structure):
    _fields_ = [("V", C.c_uint32), ("E", C.c_uint32), ("n_first", C.c_uint32), ("n_last", C.c_uint32),
                ("W", C.c_uint32), ("col", C.POINTER(C.c_uint32)), ("mask", C.POINTER(C.c_uint8)),
                ("weight", C.POINTER(C.c_float)), ("pred_off", C.POINTER(C.c_uint32)),
                ("preds", C.POINTER(C.c_uint32)), ("first", C.POINTER(C.c_uint32)), ("last", C.POINTER(C.c_uint32))]


class _Mesh(C.Structure):
    _fields_ = [("V", C.c_uint32), ("L", C.c_uint32), ("value_midx", C.POINTER(C.c_uint32)),
                ("value_sidx", C.POINTER(C.c_uint32)), ("gapm_idx", C.POINTER(C.c_uint32)),
                ("gaps_idx", C.POINTER(C.c_uint32)), ("value", C.POINTER(C.c_float)),
                ("gapm_val", C.POINTER(C.c_float)), ("gaps_val", C.POINTER(C.c_float))]


class _Index(C.Structure):
    _fields_ = [("N", C.c_uint32), ("k", C.c_int), ("nofast", C.c_int), ("n_kmers", C.c_uint64),
                ("list_off", C.POINTER(C.c_uint64)), ("postings", C.POINTER(C.c_uint32))]


_CHAR2MASK = np.zeros(256, np.int16) - 1


def _init_tables():
    for ch, m in zip("AGCTURYKMSWBDHVN", [1, 2, 4, 8, 8, 3, 12, 10, 5, 6, 9, 14, 11, 13, 7, 15]):
        _CHAR2MASK[ord(ch)] = m
        _CHAR2MASK[ord(ch.lower())] = m | 16
    _CHAR2MASK[ord("-")] = 0
    _CHAR2MASK[ord(".")] = 0


_init_tables()
MASK2RNA = np.frombuffer(b".AGRCMSVUWKDYHBN.agrcmsvuwkdyhbn", np.uint8)


def encode(seq):
    """unaligned base string -> masks (gaps dropped)"""
    a = _CHAR2MASK[np.frombuffer(seq.encode() if isinstance(seq, str) else seq, np.uint8)]
    if (a < 0).any():
        raise ValueError("bad character")
    return a[a > 0].astype(np.uint8)


def decode(masks):
    return MASK2RNA[np.asarray(masks, np.uint8) & 31].tobytes().decode()


class MSA:
    """Packed reference alignment."""

    def __init__(self, masks, cols, off, W, names=None):
        self.masks = np.ascontiguousarray(masks, np.uint8)
        self.cols = np.ascontiguousarray(cols, np.uint32)
        self.off = np.ascontiguousarray(off, np.uint64)
        self.W = int(W)
        self.N = len(self.off) - 1
        self.names = names

    @staticmethod
    def from_rows(rows, names=None):
        masks, cols, off = [], [], [0]
        W = None
        for r in rows:
            a = _CHAR2MASK[np.frombuffer(r.encode(), np.uint8)]
            if (a < 0).any():
                raise ValueError("bad character")
            W = len(a) if W is None else W
            assert len(a) == W, "rows differ in width"
            idx = np.nonzero(a > 0)[0]
            masks.append(a[idx].astype(np.uint8))
            cols.append(idx.astype(np.uint32))
            off.append(off[-1] + len(idx))
        return MSA(np.concatenate(masks) if masks else np.zeros(0, np.uint8),
                   np.concatenate(cols) if cols else np.zeros(0, np.uint32), np.array(off, np.uint64), W or 0, names)

    def row(self, i):
        a, b = int(self.off[i]), int(self.off[i + 1])
        return self.masks[a:b], self.cols[a:b]

    def row_string(self, i):
        m, c = self.row(i)
        s = np.full(self.W, ord("-"), np.uint8)
        s[c] = MASK2RNA[m & 31]
        return s.tobytes().decode()


def render(masks, cols, W):
    """cseq::getAligned(nodots=true) (src/cseq.cpp:135-174). The reference's gap placement can push the
    last bases one column past W when the right edge is crowded (its string is then longer than W);
    that quirk is reproduced, not hidden."""
    cols = np.asarray(cols, np.int64)
    n = max(W, int(cols.max()) + 1) if len(cols) else W
    s = np.full(n, ord("-"), np.uint8)
    s[cols] = MASK2RNA[np.asarray(masks, np.uint8) & 31]
    return s.tobytes().decode()


def compare_np(amasks, acols, bmasks, bcols, iupac=0, dist=0, cover=1, filter_lc=False):
    """numpy restatement of cseq_comparator::operator() (src/cseq_comparator.cpp:57-118,209-293): the merge loop's
    counters in closed form (this is the derivation the CUDA kernel uses; pinned against the compiled reference in
    tests/test_search.py). Rows are (masks, strictly increasing columns). Returns float32 (NaN for 0/0)."""
    am, ac = np.asarray(amasks, np.uint8), np.asarray(acols, np.int64)
    bm, bc = np.asarray(bmasks, np.uint8), np.asarray(bcols, np.int64)
    af = (am & 16) != 0 if filter_lc else np.zeros(len(am), bool)
    bf = (bm & 16) != 0 if filter_lc else np.zeros(len(bm), bool)
    ai, bi = np.nonzero(~af)[0], np.nonzero(~bf)[0]
    if len(ai) == 0 or len(bi) == 0:
        return np.float32(np.nan)      # the reference dereferences end() here
    ta0, ta1, tb0, tb1 = ai[0], ai[-1] + 1, bi[0], bi[-1] + 1   # filtered bases trimmed at both ends (:65-79)
    am, ac, af = am[ta0:ta1], ac[ta0:ta1], af[ta0:ta1]
    bm, bc, bf = bm[tb0:tb1], bc[tb0:tb1], bf[tb0:tb1]
    lo, hi = max(ac[0], bc[0]), min(ac[-1], bc[-1])             # columns both sequences span (:81-117)
    a_in, b_in = (ac >= lo) & (ac <= hi), (bc >= lo) & (bc <= hi)
    ovh_a, ovh_b = int((~a_in & ~af).sum()), int((~b_in & ~bf).sum())
    pos = np.searchsorted(ac, bc)
    has = (pos < len(ac)) & (ac[np.minimum(pos, len(ac) - 1)] == bc) & b_in
    pa = np.minimum(pos, len(ac) - 1)
    both_unf = has & ~bf & ~af[pa]
    x, y = am[pa] & 15, bm & 15
    if iupac == 0:
        eq = (x & y) != 0
    elif iupac == 1:
        eq = (np.array([bin(v).count("1") for v in x]) <= 1) & (x == y)
    else:
        eq = x == y
    match, mismatch = int((both_unf & eq).sum()), int((both_unf & ~eq).sum())
    only_a = int((a_in & ~af).sum()) - match - mismatch
    only_b = int((b_in & ~bf).sum()) - match - mismatch
    base = [1, match + mismatch + only_a + ovh_a, match + mismatch + only_b + ovh_b, match + mismatch + only_a + only_b,
            match + mismatch + only_a + only_b + ovh_a + ovh_b, match + mismatch + (only_a + only_b + ovh_a + ovh_b) // 2,
            match + mismatch + min(only_a + ovh_a, only_b + ovh_b), match + mismatch + max(only_a + ovh_a, only_b + ovh_b),
            match + mismatch][cover]
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.float32(match) / np.float32(base)
        if dist == 1:
            d = np.float32(-3.0 / 4 * np.log(1.0 - 4.0 / 3 * np.float64(d)))
    return np.float32(d)


class Oracle:
    """Plain-C restatement (sina_oracle.c)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.L = C.CDLL(ORACLE_SO)
        L.so_kmers.restype = C.c_int64
        L.so_kmers.argtypes = [u8p, C.c_uint32, C.c_int, C.c_int, u32p, C.c_uint64]
        L.so_index_build.restype = C.POINTER(_Index)
        L.so_index_build.argtypes = [C.c_uint32, u8p, u64p, C.c_int, C.c_int]
        L.so_index_free.argtypes = [C.POINTER(_Index)]
        L.so_turn_check.restype = C.c_int
        L.so_turn_check.argtypes = [C.POINTER(_Index), u8p, C.c_uint32, C.c_int, i32p]
        L.so_find.restype = C.c_uint32
        L.so_find.argtypes = [C.POINTER(_Index), u8p, C.c_uint32, C.c_uint32, i16p, u32p, C.POINTER(C.c_uint64)]
        L.so_family.restype = C.c_int
        L.so_family.argtypes = [C.POINTER(_Index), u64p, u32p, u8p, C.c_uint32, C.c_int64, C.POINTER(FamParams), u32p,
                                f32p, C.c_uint32, C.POINTER(C.c_uint64)]
        L.so_family_from_ranked.restype = C.c_int
        L.so_family_from_ranked.argtypes = [u32p, i16p, C.c_uint32, C.c_uint32, u64p, u32p, C.c_int64,
                                            C.POINTER(FamParams), u32p, f32p, C.c_uint32]
        L.so_graph_build.restype = C.POINTER(_Graph)
        L.so_graph_build.argtypes = [u32p, C.c_uint32, u8p, u32p, u64p, C.c_uint32, C.c_float]
        L.so_graph_free.argtypes = [C.POINTER(_Graph)]
        L.so_mesh_compute.restype = C.POINTER(_Mesh)
        L.so_mesh_compute.argtypes = [C.POINTER(_Graph), u8p, C.c_uint32, C.POINTER(AlignParams)]
        L.so_mesh_free.argtypes = [C.POINTER(_Mesh)]
        L.so_backtrack.restype = C.c_int
        L.so_backtrack.argtypes = [C.POINTER(_Graph), C.POINTER(_Mesh), u8p, C.c_uint32, C.POINTER(AlignParams),
                                   C.POINTER(OracleResult), u32p, u8p]
        L.so_fix_duplicate_positions.restype = C.c_int
        L.so_fix_duplicate_positions.argtypes = [u32p, u8p, C.c_uint32, C.c_uint32, C.c_int]
        L.so_align.restype = C.c_int
        L.so_align.argtypes = [u32p, C.c_uint32, u8p, u32p, u64p, C.c_uint32, u8p, C.c_uint32, C.POINTER(AlignParams),
                               C.POINTER(OracleResult), u32p, u8p]
        L.so_run_batch.restype = C.c_int
        L.so_run_batch.argtypes = [C.POINTER(_Index), u8p, u32p, u64p, C.c_uint32, C.c_uint32, u8p, u64p, C.c_void_p,
                                   C.POINTER(FamParams), C.POINTER(AlignParams), C.c_int, C.POINTER(OracleResult),
                                   u32p, u8p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.so_char_to_mask.restype = C.c_int
        L.so_mask_to_char.restype = C.c_int
        L.so_set_column_weights.argtypes = [C.c_void_p, C.c_uint32]
        self._colw = None

    def set_column_weights(self, w):
        """positional weights (scoring_scheme_weighted) for every later mesh / align / run_batch call; None = off"""
        self._colw = None if w is None else np.ascontiguousarray(w, np.float32)
        if self._colw is None:
            self.L.so_set_column_weights(None, 0)
        else:
            self.L.so_set_column_weights(self._colw.ctypes.data_as(C.c_void_p), len(self._colw))

    def char_to_mask(self, c):
        return self.L.so_char_to_mask(C.c_int(ord(c)))

    def mask_to_char(self, m, dna=False):
        return chr(self.L.so_mask_to_char(C.c_int(m), C.c_int(int(dna))))

    def kmers(self, masks, k, mode):
        masks = np.ascontiguousarray(masks, np.uint8)
        out = np.zeros(max(1, len(masks)), np.uint32)
        n = self.L.so_kmers(masks if len(masks) else np.zeros(1, np.uint8), len(masks), k, mode, out, len(out))
        return out[:n].copy()

    def index_build(self, msa, k=10, nofast=False):
        return self.L.so_index_build(msa.N, msa.masks, msa.off, k, int(nofast))

    def index_free(self, ix):
        self.L.so_index_free(ix)

    def index_lists(self, ix):
        """(list_off, postings) numpy copies"""
        c = ix.contents
        off = np.ctypeslib.as_array(c.list_off, (c.n_kmers + 1,)).copy()
        post = np.ctypeslib.as_array(c.postings, (max(1, int(off[-1])),)).copy()[:int(off[-1])]
        return off, post

    def find(self, ix, q, max_results):
        q = np.ascontiguousarray(q, np.uint8)
        n = min(max_results, ix.contents.N)
        sc = np.zeros(max(1, n), np.int16)
        ids = np.zeros(max(1, n), np.uint32)
        P = C.c_uint64()
        r = self.L.so_find(ix, q, len(q), max_results, sc, ids, C.byref(P))
        return sc[:r].copy(), ids[:r].copy(), P.value

    def turn_check(self, ix, q, all_frames=True):
        """famfinder::turn_check: (best orientation 0..3, the four top scores)"""
        q = np.ascontiguousarray(q, np.uint8)
        sc = np.zeros(4, np.int32)
        return self.L.so_turn_check(ix, q, len(q), int(all_frames), sc), sc

    def family(self, ix, msa, q, fp=None, exclude_id=-1):
        fp = fp or FamParams()
        q = np.ascontiguousarray(q, np.uint8)
        cap = fp.fs_max * 4 + 64
        ids = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.float32)
        P = C.c_uint64()
        n = self.L.so_family(ix, msa.off, msa.cols, q, len(q), exclude_id, C.byref(fp), ids, sc, cap, C.byref(P))
        return n, ids[:max(n, 0)].copy(), sc[:max(n, 0)].copy()

    def family_from_ranked(self, cand_ids, cand_scores, n_total, msa, fp=None, exclude_id=-1):
        fp = fp or FamParams()
        cap = fp.fs_max * 4 + 64
        ids = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.float32)
        cand_ids = np.ascontiguousarray(cand_ids, np.uint32)
        cand_scores = np.ascontiguousarray(cand_scores, np.int16)
        n = self.L.so_family_from_ranked(cand_ids, cand_scores, len(cand_ids), n_total, msa.off, msa.cols,
                                         exclude_id, C.byref(fp), ids, sc, cap)
        return n, ids[:max(n, 0)].copy(), sc[:max(n, 0)].copy()

    def graph(self, msa, fam, fs_weight=1.0):
        fam = np.ascontiguousarray(fam, np.uint32)
        g = self.L.so_graph_build(fam, len(fam), msa.masks, msa.cols, msa.off, msa.W, fs_weight)
        c = g.contents
        V, E = c.V, c.E
        arr = lambda p, n, dt: np.ctypeslib.as_array(p, (max(n, 1),)).astype(dt)[:n].copy()
        d = dict(V=V, E=E, col=arr(c.col, V, np.uint32), mask=arr(c.mask, V, np.uint8),
                 weight=arr(c.weight, V, np.float32), pred_off=arr(c.pred_off, V + 1, np.uint32),
                 preds=arr(c.preds, E, np.uint32), first=arr(c.first, c.n_first, np.uint32),
                 last=arr(c.last, c.n_last, np.uint32))
        self.L.so_graph_free(g)
        return d

    def mesh(self, msa, fam, q, ap=None):
        """full mesh as dict of (V, L) arrays, plus graph size"""
        ap = ap or AlignParams()
        fam = np.ascontiguousarray(fam, np.uint32)
        q = np.ascontiguousarray(q, np.uint8)
        g = self.L.so_graph_build(fam, len(fam), msa.masks, msa.cols, msa.off, msa.W, ap.fs_weight)
        m = self.L.so_mesh_compute(g, q, len(q), C.byref(ap))
        V, Lq = g.contents.V, len(q)
        out = {}
        for name, dt in [("value_midx", np.uint32), ("value_sidx", np.uint32), ("gapm_idx", np.uint32),
                         ("gaps_idx", np.uint32), ("value", np.float32), ("gapm_val", np.float32),
                         ("gaps_val", np.float32)]:
            out[name] = np.ctypeslib.as_array(getattr(m.contents, name), (V * Lq,)).astype(dt).reshape(V, Lq).copy()
        self.L.so_mesh_free(m)
        self.L.so_graph_free(g)
        return out

    def fix_duplicate_positions(self, pos, masks, width, lowercase=False):
        pos = np.array(pos, np.uint32)
        masks = np.array(masks, np.uint8)
        st = self.L.so_fix_duplicate_positions(pos if len(pos) else np.zeros(1, np.uint32),
                                               masks if len(masks) else np.zeros(1, np.uint8), len(pos), width,
                                               int(lowercase))
        return st, pos, masks

    def align(self, msa, fam, q, ap=None):
        """returns (result struct, out_cols, out_masks, permuted family)"""
        ap = ap or AlignParams()
        fam = np.array(fam, np.uint32)
        q = np.ascontiguousarray(q, np.uint8)
        r = OracleResult()
        oc = np.zeros(max(1, len(q)), np.uint32)
        om = np.zeros(max(1, len(q)), np.uint8)
        self.L.so_align(fam, len(fam), msa.masks, msa.cols, msa.off, msa.W, q, len(q), C.byref(ap), C.byref(r), oc, om)
        return r, oc[:r.n_out].copy(), om[:r.n_out].copy(), fam

    def run_batch(self, ix, msa, qmasks, qoff, fp=None, ap=None, nthreads=0, exclude_ids=None):
        fp = fp or FamParams()
        ap = ap or AlignParams()
        nq = len(qoff) - 1
        qmasks = np.ascontiguousarray(qmasks, np.uint8)
        qoff = np.ascontiguousarray(qoff, np.uint64)
        res = (OracleResult * nq)()
        oc = np.zeros(max(1, len(qmasks)), np.uint32)
        om = np.zeros(max(1, len(qmasks)), np.uint8)
        cells, posts = C.c_uint64(), C.c_uint64()
        ex = None
        if exclude_ids is not None:
            ex_arr = np.ascontiguousarray(exclude_ids, np.int64)
            ex = ex_arr.ctypes.data_as(C.c_void_p)
        nt = self.L.so_run_batch(ix, msa.masks, msa.cols, msa.off, msa.W, nq, qmasks, qoff, ex, C.byref(fp),
                                 C.byref(ap), nthreads, res, oc, om, C.byref(cells), C.byref(posts))
        return res, oc, om, cells.value, posts.value, nt


class Ref:
    """The reference's own code (libsina_ref.so). Raises FileNotFoundError if the prebuilt .so is absent."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            if os.path.isdir("/root/reference/src"):
                build(ref=True)
            else:
                raise FileNotFoundError(REF_SO)
        L = self.L = C.CDLL(REF_SO)
        L.ref_db_create.restype = C.c_void_p
        L.ref_db_create_packed.restype = C.c_void_p
        L.ref_db_create_packed.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, u8p, u32p, u64p]
        L.ref_db_free.argtypes = [C.c_void_p]
        L.ref_kidx_build.restype = C.c_void_p
        L.ref_kidx_build.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_kidx_free.argtypes = [C.c_void_p]
        L.ref_kidx_store.restype = C.c_int
        L.ref_kidx_store.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        L.ref_kidx_from_lists.restype = C.c_void_p
        L.ref_kidx_from_lists.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_uint32, u32p, u64p, u32p]
        L.ref_kidx_list_scores.restype = C.c_int
        L.ref_kidx_list_scores.argtypes = [C.c_void_p, C.c_uint32, i16p]
        L.ref_kidx_load.restype = C.c_void_p
        L.ref_kidx_load.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.ref_kidx_list_size.restype = C.c_uint64
        L.ref_kidx_list_size.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_kidx_find.restype = C.c_uint32
        L.ref_kidx_find.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, i16p, u32p, C.POINTER(C.c_uint64)]
        L.ref_turn_check.restype = C.c_int
        L.ref_turn_check.argtypes = [C.c_void_p, C.c_char_p, C.c_int, i32p]
        L.ref_family.restype = C.c_int
        L.ref_family.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(FamParams), u32p, f32p, C.c_uint32]
        L.ref_graph.restype = C.c_int
        L.ref_graph.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, u32p, u8p, f32p, u32p,
                                u32p, C.POINTER(C.c_uint32), u32p, C.POINTER(C.c_uint32), u32p,
                                C.POINTER(C.c_uint32)]
        L.ref_align.restype = C.c_int
        L.ref_align.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_char_p, C.c_char_p, C.POINTER(AlignParams),
                                C.POINTER(RefResult), C.c_char_p, u32p, C.POINTER(C.c_uint32), C.c_char_p,
                                C.c_uint32, C.c_void_p, C.c_uint64]
        L.ref_kmers.restype = C.c_int
        L.ref_kmers.argtypes = [C.c_char_p, C.c_int, C.c_int, u32p, C.c_int]
        L.ref_compare.restype = C.c_float
        L.ref_compare.argtypes = [C.c_uint32, u8p, u32p, C.c_uint32, u8p, u32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_search.restype = C.c_int
        L.ref_search.argtypes = [C.c_void_p, C.c_uint32, u8p, u32p, C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, u32p, f32p]
        L.ref_vlimap_increment.restype = C.c_int
        L.ref_vlimap_increment.argtypes = [C.c_uint32, u32p, C.c_uint32, C.c_int, i16p]
        L.ref_fix_duplicate_positions.restype = C.c_int
        L.ref_fix_duplicate_positions.argtypes = [C.c_uint32, u32p, u8p, C.c_uint32, C.c_int, u32p, u8p]
        L.ref_char_to_mask.restype = C.c_int
        L.ref_set_column_weights.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.ref_cseq_roundtrip.restype = C.c_int
        L.ref_run_batch.restype = C.c_int
        L.ref_run_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(FamParams),
                                    C.POINTER(AlignParams), C.c_int, u64p, u32p, C.POINTER(RefResult),
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]

    def set_column_weights(self, w, pad=1 << 16):
        """alignment_stats weights (--filter) for every later align / run_batch call; None = scoring_scheme_simple"""
        if w is None:
            self.L.ref_set_column_weights(None, 0, 0)
        else:
            w = np.ascontiguousarray(w, np.float32)
            self.L.ref_set_column_weights(w.ctypes.data_as(C.c_void_p), len(w), pad)

    def char_to_mask(self, c):
        return self.L.ref_char_to_mask(C.c_ubyte(ord(c)))

    def cseq_roundtrip(self, row, nodots=True, dna=False):
        buf = C.create_string_buffer(len(row) + 8)
        n = self.L.ref_cseq_roundtrip(row.encode(), int(nodots), int(dna), buf, len(row) + 8)
        return buf.value.decode() if n >= 0 else None

    def db(self, msa):
        chars = MASK2RNA[msa.masks & 31].copy()
        names = None
        if msa.names is not None:
            arr = (C.c_char_p * msa.N)(*[n.encode() for n in msa.names])
            names = C.cast(arr, C.c_void_p)
        return C.c_void_p(self.L.ref_db_create_packed(msa.N, msa.W, names, chars, msa.cols, msa.off))

    def db_free(self, db):
        self.L.ref_db_free(db)

    def kmers(self, seq, k, mode):
        out = np.zeros(max(1, len(seq)), np.uint32)
        n = self.L.ref_kmers(seq.encode(), k, mode, out, len(out))
        return out[:n].copy()

    def vlimap_increment(self, maxsize, ids, invert):
        ids = np.ascontiguousarray(ids, np.uint32)
        sc = np.zeros(max(1, maxsize), np.int16)
        r = self.L.ref_vlimap_increment(maxsize, ids if len(ids) else np.zeros(1, np.uint32), len(ids), int(invert), sc)
        return r, sc[:maxsize]

    def kidx_build(self, db, k=10, nofast=False):
        return C.c_void_p(self.L.ref_kidx_build(db, k, int(nofast)))

    def kidx_free(self, ix):
        self.L.ref_kidx_free(ix)

    def kidx_from_lists(self, N, k, nofast, kmers, list_off, ids):
        """index made of the given posting lists (reference vlimap objects, inverted when longer than N / 2)"""
        kmers = np.ascontiguousarray(kmers, np.uint32)
        ids = np.ascontiguousarray(ids if len(ids) else np.zeros(1), np.uint32)
        return C.c_void_p(self.L.ref_kidx_from_lists(N, k, int(nofast), len(kmers), kmers, np.ascontiguousarray(list_off, np.uint64), ids))

    def kidx_list_scores(self, ix, kmer, N):
        """(scores[N], offset) of one list's vlimap::increment, offset added: 1 where the id is in the list"""
        sc = np.zeros(max(1, N), np.int16)
        off = self.L.ref_kidx_list_scores(ix, kmer, sc)
        return sc[:N], off

    def kidx_store(self, ix, names, path):
        """kmer_search::impl::store: the reference's .sidx file for this index"""
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        assert self.L.ref_kidx_store(ix, C.cast(arr, C.c_void_p), str(path).encode()) == 0

    def kidx_load(self, db, path, k=10, nofast=False):
        """kmer_search::impl::try_load: index from a .sidx file (None on a header mismatch)"""
        h = self.L.ref_kidx_load(db, str(path).encode(), k, int(nofast))
        return C.c_void_p(h) if h else None

    def kidx_list_size(self, ix, kmer):
        return self.L.ref_kidx_list_size(ix, kmer)

    def find(self, ix, query, max_results):
        sc = np.zeros(max(1, max_results), np.int16)
        ids = np.zeros(max(1, max_results), np.uint32)
        P = C.c_uint64()
        n = self.L.ref_kidx_find(ix, query.encode(), max_results, sc, ids, C.byref(P))
        return sc[:n].copy(), ids[:n].copy(), P.value

    def compare(self, amasks, acols, bmasks, bcols, iupac=0, dist=0, cover=1, filter_lc=False):
        """cseq_comparator::operator() (src/cseq_comparator.cpp:209-293) on two aligned rows given as (masks, columns)"""
        ac, bc = MASK2RNA[np.asarray(amasks, np.uint8) & 31].copy(), MASK2RNA[np.asarray(bmasks, np.uint8) & 31].copy()
        return float(self.L.ref_compare(len(ac), ac, np.ascontiguousarray(acols, np.uint32), len(bc), bc,
                                        np.ascontiguousarray(bcols, np.uint32), iupac, dist, cover, int(filter_lc)))

    def search(self, ix, masks, cols, kmer_candidates=1000, max_result=10, min_sim=0.7, ignore_super=False, iupac=0,
               dist=0, cover=1, filter_lc=False):
        """search_filter::operator() (src/search_filter.cpp:244-330, k-mer branch) for one aligned query"""
        ch = MASK2RNA[np.asarray(masks, np.uint8) & 31].copy()
        ids, sc = np.zeros(max(1, max_result), np.uint32), np.zeros(max(1, max_result), np.float32)
        n = self.L.ref_search(ix, len(ch), ch, np.ascontiguousarray(cols, np.uint32), kmer_candidates, max_result,
                              min_sim, int(ignore_super), iupac, dist, cover, int(filter_lc), ids, sc)
        return ids[:n].copy(), sc[:n].copy()

    def turn_check(self, ix, query, all_frames=True):
        sc = np.zeros(4, np.int32)
        return self.L.ref_turn_check(ix, query.encode(), int(all_frames), sc), sc

    def family(self, ix, query, fp=None, qname=""):
        fp = fp or FamParams()
        cap = fp.fs_max * 4 + 64
        ids = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.float32)
        n = self.L.ref_family(ix, qname.encode(), query.encode(), C.byref(fp), ids, sc, cap)
        return n, ids[:max(n, 0)].copy(), sc[:max(n, 0)].copy()

    def graph(self, db, fam, fs_weight=1.0, cap_nodes=1 << 16, cap_edges=1 << 18):
        fam = np.ascontiguousarray(fam, np.uint32)
        col = np.zeros(cap_nodes, np.uint32); ch = np.zeros(cap_nodes, np.uint8); w = np.zeros(cap_nodes, np.float32)
        po = np.zeros(cap_nodes + 1, np.uint32); pr = np.zeros(cap_edges, np.uint32)
        first = np.zeros(cap_nodes, np.uint32); last = np.zeros(cap_nodes, np.uint32)
        ne, nf, nl = C.c_uint32(), C.c_uint32(), C.c_uint32()
        V = self.L.ref_graph(db, fam, len(fam), fs_weight, cap_nodes, cap_edges, col, ch, w, po, pr, C.byref(ne), first,
                             C.byref(nf), last, C.byref(nl))
        assert V >= 0
        mask = _CHAR2MASK[ch[:V]].astype(np.uint8)
        return dict(V=V, E=ne.value, col=col[:V].copy(), mask=mask, weight=w[:V].copy(), pred_off=po[:V + 1].copy(),
                    preds=pr[:ne.value].copy(), first=first[:nf.value].copy(), last=last[:nl.value].copy())

    def align(self, db, fam, query, W, ap=None, want_cells=False, qname="q"):
        """returns (RefResult, aligned string, cols, log, cells dict or None)"""
        ap = ap or AlignParams()
        fam = np.ascontiguousarray(fam, np.uint32)
        r = RefResult()
        out = C.create_string_buffer(W + 8)
        nb = len(encode(query))
        oc = np.zeros(max(1, nb), np.uint32)
        ncols = C.c_uint32()
        log = C.create_string_buffer(4096)
        cells = None
        cells_p, cap = None, 0
        if want_cells:
            g = self.graph(db, fam, ap.fs_weight)
            cap = g["V"] * nb * 7
            cells = np.zeros(max(1, cap), np.uint32)
            cells_p = cells.ctypes.data_as(C.c_void_p)
        self.L.ref_align(db, fam, len(fam), qname.encode(), query.encode(), C.byref(ap), C.byref(r), out, oc,
                         C.byref(ncols), log, 4096, cells_p, cap)
        cd = None
        if want_cells and r.status in (0, 3):
            c = cells[:r.n_nodes * nb * 7].reshape(-1, nb, 7)
            cd = dict(value_midx=c[:, :, 0].copy(), value_sidx=c[:, :, 1].copy(), gapm_idx=c[:, :, 2].copy(),
                      gaps_idx=c[:, :, 3].copy(), value=c[:, :, 4].copy().view(np.float32),
                      gapm_val=c[:, :, 5].copy().view(np.float32), gaps_val=c[:, :, 6].copy().view(np.float32))
        return r, out.value.decode(), oc[:ncols.value].copy(), log.value.decode(), cd

    def run_batch(self, ix, queries, fp=None, ap=None, nthreads=0, names=None):
        """whole path (family + align) over strings; returns (results, cols, qoff, cells, postings, threads)"""
        fp = fp or FamParams()
        ap = ap or AlignParams()
        nq = len(queries)
        qarr = (C.c_char_p * nq)(*[q.encode() for q in queries])
        narr = None
        if names is not None:
            narr_ = (C.c_char_p * nq)(*[n.encode() for n in names])
            narr = C.cast(narr_, C.c_void_p)
        lens = np.array([len(encode(q)) for q in queries], np.uint64)
        qoff = np.zeros(nq + 1, np.uint64)
        qoff[1:] = np.cumsum(lens)
        oc = np.zeros(max(1, int(qoff[-1])), np.uint32)
        res = (RefResult * nq)()
        cells, posts = C.c_uint64(), C.c_uint64()
        nt = self.L.ref_run_batch(ix, nq, narr, C.cast(qarr, C.c_void_p), C.byref(fp), C.byref(ap), nthreads, qoff, oc,
                                  res, C.byref(cells), C.byref(posts))
        return res, oc, qoff, cells.value, posts.value, nt


def have_ref():
    return os.path.exists(REF_SO) or os.path.isdir("/root/reference/src")
