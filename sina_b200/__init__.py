"""sina_b200 -- B200-native replacement for SINA's per-query hot path (k-mer family finding, family
graph, mesh DP, backtrack, gap placement) behind a C-ABI (include/sina_b200.h).

This module is the thin ctypes view of sina_b200/libsina_b200.so used by the tests and bench.py. The
compute path is the CUDA library only: importing works without a GPU (so that symbol checks can run on a
CPU box), but every compute call raises if the library or a CUDA device is missing -- there is no CPU
fallback and nothing here imports oracle/.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get("SINA_B200_LIB", "libsina_b200.so"))   # SINA_B200_LIB: an experimental build next to the product library (tools/)

# every symbol include/sina_b200.h declares
EXPORTS = [
    "sg_default_fam_params", "sg_default_align_params", "sg_last_error", "sg_device_count",
    "sg_index_create", "sg_index_destroy", "sg_index_info", "sg_index_list_sizes", "sg_index_list", "sg_index_set_column_weights", "sg_index_export_lists",
    "sg_find_batch", "sg_turn_batch", "sg_family_batch", "sg_align_batch", "sg_run_batch",
    "sg_default_search_params", "sg_index_set_name_ranks", "sg_identity_batch", "sg_search_batch",
    "sg_family_batch_aligned", "sg_session_set_query_columns",
    "sg_session_create", "sg_session_destroy", "sg_session_upload", "sg_session_find", "sg_session_turn", "sg_session_family",
    "sg_session_set_family", "sg_session_align", "sg_session_run", "sg_session_sync", "sg_session_download_find",
    "sg_session_download_family", "sg_session_download_align", "sg_session_stats", "sg_session_timer", "sg_session_dump_graph",
]

SG_Q_ALIGNED, SG_Q_COPIED, SG_Q_SKIPPED, SG_Q_NOSPACE, SG_Q_NOFAMILY, SG_Q_LIMIT = 0, 1, 2, 3, 4, 5
TURN_MODES = {"none": 0, "revcomp": 1, "all": 2}

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C")


class FamParams(C.Structure):
    """sg_fam_params; defaults are SINA's (src/famfinder.cpp:155-195)."""
    _fields_ = [("fs_min", C.c_uint32), ("fs_max", C.c_uint32), ("fs_msc", C.c_float), ("fs_msc_max", C.c_float),
                ("fs_min_len", C.c_uint32), ("fs_req_full", C.c_uint32), ("fs_full_len", C.c_uint32),
                ("fs_req_gaps", C.c_uint32), ("fs_req", C.c_uint32), ("leave_query_out", C.c_int32)]

    def __init__(self, fs_min=40, fs_max=40, fs_msc=0.7, fs_msc_max=2.0, fs_min_len=150, fs_req_full=1,
                 fs_full_len=1400, fs_req_gaps=10, fs_req=1, leave_query_out=0):
        super().__init__(fs_min, fs_max, fs_msc, fs_msc_max, fs_min_len, fs_req_full, fs_full_len, fs_req_gaps,
                         fs_req, leave_query_out)


class AlignParams(C.Structure):
    """sg_align_params; defaults are SINA's (src/align.cpp:232-259)."""
    _fields_ = [("match_score", C.c_float), ("mismatch_score", C.c_float), ("gap_penalty", C.c_float),
                ("gap_ext_penalty", C.c_float), ("fs_weight", C.c_float), ("overhang", C.c_int32),
                ("lowercase", C.c_int32), ("insertion", C.c_int32), ("realign", C.c_int32)]

    def __init__(self, match_score=2.0, mismatch_score=-1.0, gap_penalty=5.0, gap_ext_penalty=2.0, fs_weight=1.0,
                 overhang=0, lowercase=0, insertion=0, realign=0):
        super().__init__(match_score, mismatch_score, gap_penalty, gap_ext_penalty, fs_weight, overhang, lowercase,
                         insertion, realign)


IUPAC_RULES = {"optimistic": 0, "pessimistic": 1, "exact": 2}
CORRECTIONS = {"none": 0, "jc": 1}
COVER_RULES = {"abs": 0, "query": 1, "target": 2, "overlap": 3, "all": 4, "average": 5, "min": 6, "max": 7, "nogap": 8}


class SearchParams(C.Structure):
    """sg_search_params; defaults are SINA's (src/search_filter.cpp:96-133, src/cseq_comparator.cpp:432-462)."""
    _fields_ = [("kmer_candidates", C.c_uint32), ("max_result", C.c_uint32), ("min_sim", C.c_float),
                ("ignore_super", C.c_int32), ("iupac", C.c_int32), ("correction", C.c_int32), ("cover", C.c_int32),
                ("filter_lowercase", C.c_int32)]

    def __init__(self, kmer_candidates=1000, max_result=10, min_sim=0.7, ignore_super=0, iupac=0, correction=0, cover=1,
                 filter_lowercase=0):
        super().__init__(kmer_candidates, max_result, min_sim, ignore_super, iupac, correction, cover, filter_lowercase)


class AlignResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("score", C.c_float), ("raw", C.c_float), ("sum_weight", C.c_float),
                ("head", C.c_int32), ("tail", C.c_int32), ("qual", C.c_int32), ("n_nodes", C.c_uint32),
                ("fam_used", C.c_uint32), ("n_out", C.c_uint32), ("end_m", C.c_uint32), ("end_s", C.c_uint32)]


RESULT_DTYPE = np.dtype([("status", np.int32), ("score", np.float32), ("raw", np.float32), ("sum_weight", np.float32),
                         ("head", np.int32), ("tail", np.int32), ("qual", np.int32), ("n_nodes", np.uint32),
                         ("fam_used", np.uint32), ("n_out", np.uint32), ("end_m", np.uint32), ("end_s", np.uint32)])


class StageStats(C.Structure):
    _fields_ = [("ms_find", C.c_float), ("ms_family", C.c_float), ("ms_graph", C.c_float), ("ms_dp", C.c_float),
                ("ms_backtrack", C.c_float), ("cells", C.c_uint64), ("postings", C.c_uint64),
                ("kernel_launches", C.c_uint64)]


class SinaB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Load libsina_b200.so (fails loudly if it was not built: `make -C sina_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SinaB200Error("%s is missing: build it with `make -C sina_b200/csrc` (or __graft_entry__.build()); "
                            "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.sg_last_error.restype = C.c_char_p
    L.sg_index_create.argtypes = [u8p, u32p, u64p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_void_p)]
    L.sg_index_destroy.argtypes = [C.c_void_p]
    L.sg_index_destroy.restype = None
    L.sg_index_info.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    L.sg_index_set_column_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.sg_index_export_lists.argtypes = [C.c_void_p, u64p, C.c_void_p]
    L.sg_index_list.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.sg_index_list_sizes.argtypes = [C.c_void_p, u32p, C.c_uint32, u64p]
    L.sg_find_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint32, C.c_uint32, i16p, u32p, u32p]
    L.sg_turn_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint32, C.c_int, i32p]
    L.sg_family_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint32, C.c_void_p, C.POINTER(FamParams), C.c_uint32,
                                  u32p, f32p, i32p]
    L.sg_family_batch_aligned.argtypes = [C.c_void_p, u8p, C.c_void_p, u64p, C.c_uint32, C.c_void_p, C.POINTER(FamParams),
                                          C.c_uint32, u32p, f32p, i32p]
    L.sg_session_set_query_columns.argtypes = [C.c_void_p, u32p]
    L.sg_align_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint32, u32p, u64p, C.POINTER(AlignParams), u32p, u8p,
                                 C.c_void_p]
    L.sg_run_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_uint32, C.c_void_p, C.POINTER(FamParams),
                               C.POINTER(AlignParams), u32p, u8p, C.c_void_p]
    L.sg_default_search_params.argtypes = [C.POINTER(SearchParams)]
    L.sg_default_search_params.restype = None
    L.sg_index_set_name_ranks.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.sg_identity_batch.argtypes = [C.c_void_p, u8p, u32p, u64p, C.c_uint32, u32p, u64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    f32p]
    L.sg_search_batch.argtypes = [C.c_void_p, u8p, u32p, u64p, C.c_uint32, C.POINTER(SearchParams), u32p, f32p, u32p]
    L.sg_session_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]
    L.sg_session_destroy.argtypes = [C.c_void_p]
    L.sg_session_destroy.restype = None
    L.sg_session_upload.argtypes = [C.c_void_p, u8p, u64p, C.c_uint32, C.c_void_p]
    L.sg_session_find.argtypes = [C.c_void_p, C.c_uint32]
    L.sg_session_turn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.sg_session_family.argtypes = [C.c_void_p, C.POINTER(FamParams)]
    L.sg_session_set_family.argtypes = [C.c_void_p, u32p, u64p]
    L.sg_session_run.argtypes = [C.c_void_p, C.POINTER(FamParams), C.POINTER(AlignParams)]
    L.sg_session_align.argtypes = [C.c_void_p, C.POINTER(AlignParams)]
    L.sg_session_sync.argtypes = [C.c_void_p]
    L.sg_session_download_find.argtypes = [C.c_void_p, i16p, u32p, u32p]
    L.sg_session_download_family.argtypes = [C.c_void_p, C.c_uint32, u32p, f32p, i32p]
    L.sg_session_download_align.argtypes = [C.c_void_p, u32p, u8p, C.c_void_p]
    L.sg_session_stats.argtypes = [C.c_void_p, C.POINTER(StageStats), C.c_int]
    L.sg_session_timer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    L.sg_session_dump_graph.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint32), u32p, u8p, f32p, u32p, u32p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise SinaB200Error("sina_b200 error %d: %s" % (rc, lib().sg_last_error().decode()))


def device_count():
    return lib().sg_device_count()


def _excl_ptr(exclude_ids):
    if exclude_ids is None:
        return None, None
    arr = np.ascontiguousarray(exclude_ids, np.int64)
    return arr, arr.ctypes.data_as(C.c_void_p)


class Index:
    """Reference MSA + k-mer posting lists resident on one GPU (sg_index). Replaces
    kmer_search::get_kmer_search(db, k, nofast) (src/kmer_search.cpp:118-134)."""

    def __init__(self, masks, cols, off, W, k=10, nofast=False, device=0):
        self.masks = np.ascontiguousarray(masks, np.uint8)
        self.cols = np.ascontiguousarray(cols, np.uint32)
        self.off = np.ascontiguousarray(off, np.uint64)
        self.N, self.W, self.k, self.nofast = len(self.off) - 1, int(W), int(k), bool(nofast)
        self.h = C.c_void_p()
        _check(lib().sg_index_create(self.masks, self.cols, self.off, self.N, self.W, self.k, int(self.nofast),
                                     device, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().sg_index_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export_lists(self):
        """every posting list: (list_off[n_slots + 1], ids[n_postings]), ids ascending inside a list"""
        inf = self.info()
        n_slots = 4 ** (inf["k"] if inf["nofast"] else inf["k"] - 1)
        off = np.zeros(n_slots + 1, np.uint64)
        ids = np.zeros(max(1, inf["n_postings"]), np.uint32)
        _check(lib().sg_index_export_lists(self.h, off, ids.ctypes.data_as(C.c_void_p)))
        return off, ids[:inf["n_postings"]]

    def set_column_weights(self, w):
        """positional column weights (--filter / alignment_stats): scoring_scheme_weighted from now on; None = off"""
        if w is None:
            _check(lib().sg_index_set_column_weights(self.h, None, 0))
        else:
            w = np.ascontiguousarray(w, np.float32)
            _check(lib().sg_index_set_column_weights(self.h, w.ctypes.data_as(C.c_void_p), len(w)))

    def set_name_ranks(self, names):
        """order of the references' names: search::result_item breaks score ties by name (src/search.h:56-68). `names` =
        list of N strings (or None: ties go by reference id)"""
        if names is None:
            _check(lib().sg_index_set_name_ranks(self.h, None, 0))
            return
        order = sorted(range(len(names)), key=lambda i: (names[i], i))
        rank = np.zeros(len(names), np.uint32)
        rank[np.asarray(order)] = np.arange(len(names), dtype=np.uint32)
        _check(lib().sg_index_set_name_ranks(self.h, rank.ctypes.data_as(C.c_void_p), len(rank)))

    def identity(self, amasks, acols, aoff, ref_ids, ref_off, iupac=0, correction=0, cover=1, filter_lowercase=0):
        """cseq_comparator::operator() of aligned sequences against reference rows (flat pair list per query)"""
        amasks, acols = np.ascontiguousarray(amasks, np.uint8), np.ascontiguousarray(acols, np.uint32)
        aoff, ref_off = np.ascontiguousarray(aoff, np.uint64), np.ascontiguousarray(ref_off, np.uint64)
        ref_ids = np.ascontiguousarray(ref_ids, np.uint32)
        out = np.zeros(max(1, len(ref_ids)), np.float32)
        _check(lib().sg_identity_batch(self.h, amasks, acols, aoff, len(aoff) - 1, ref_ids if len(ref_ids) else np.zeros(1, np.uint32),
                                       ref_off, iupac, correction, cover, filter_lowercase, out))
        return out[:len(ref_ids)]

    def search(self, amasks, acols, aoff, sp=None):
        """--search stage (search_filter::operator()) for a batch of aligned sequences: (ids[nq,max_result],
        scores[nq,max_result], n[nq])"""
        sp = sp or SearchParams()
        amasks, acols = np.ascontiguousarray(amasks, np.uint8), np.ascontiguousarray(acols, np.uint32)
        aoff = np.ascontiguousarray(aoff, np.uint64)
        nq = len(aoff) - 1
        ids, sc, n = np.zeros((nq, sp.max_result), np.uint32), np.zeros((nq, sp.max_result), np.float32), np.zeros(nq, np.uint32)
        _check(lib().sg_search_batch(self.h, amasks, acols, aoff, nq, C.byref(sp), ids, sc, n))
        return ids, sc, n

    def info(self):
        N, W, nt, ts = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        k, nf = C.c_int(), C.c_int()
        P = C.c_uint64()
        _check(lib().sg_index_info(self.h, C.byref(N), C.byref(W), C.byref(k), C.byref(nf), C.byref(P), C.byref(nt),
                                   C.byref(ts)))
        return dict(N=N.value, W=W.value, k=k.value, nofast=nf.value, n_postings=P.value, n_tiles=nt.value,
                    tile_size=ts.value)

    def list_sizes(self, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint32)
        out = np.zeros(len(kmers), np.uint64)
        _check(lib().sg_index_list_sizes(self.h, kmers, len(kmers), out))
        return out

    def posting_list(self, kmer):
        n = C.c_uint64()
        _check(lib().sg_index_list(self.h, int(kmer), None, 0, C.byref(n)))
        ids = np.zeros(max(1, n.value), np.uint32)
        _check(lib().sg_index_list(self.h, int(kmer), ids.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return ids[:n.value]

    # ---- host-buffer entry points (one call = upload + kernels + download)
    def find(self, qmasks, qoff, max_results):
        """search::find for a batch: (scores[nq,max], ids[nq,max], nres[nq]) in rank order."""
        qmasks, qoff = np.ascontiguousarray(qmasks, np.uint8), np.ascontiguousarray(qoff, np.uint64)
        nq, m = len(qoff) - 1, min(max_results, self.N)
        sc, ids, nres = np.zeros((nq, m), np.int16), np.zeros((nq, m), np.uint32), np.zeros(nq, np.uint32)
        _check(lib().sg_find_batch(self.h, qmasks, qoff, nq, max_results, sc, ids, nres))
        return sc, ids, nres

    def turn(self, qmasks, qoff, mode="all"):
        """--turn orientation check (famfinder::turn_check): 0 none, 1 reversed, 2 complemented, 3 both, per query."""
        qmasks, qoff = np.ascontiguousarray(qmasks, np.uint8), np.ascontiguousarray(qoff, np.uint64)
        nq = len(qoff) - 1
        out = np.zeros(nq, np.int32)
        _check(lib().sg_turn_batch(self.h, qmasks, qoff, nq, TURN_MODES[mode], out))
        return out

    def family(self, qmasks, qoff, fp=None, exclude_ids=None, qcols=None):
        """famfinder stage; `qcols` = positions of the query bases (pre-aligned input), needed by fs_msc_max < 1"""
        fp = fp or FamParams()
        qmasks, qoff = np.ascontiguousarray(qmasks, np.uint8), np.ascontiguousarray(qoff, np.uint64)
        nq, stride = len(qoff) - 1, max(fp.fs_min, fp.fs_max) + fp.fs_req_full + 1
        ids, sc, n = np.zeros((nq, stride), np.uint32), np.zeros((nq, stride), np.float32), np.zeros(nq, np.int32)
        keep, ex = _excl_ptr(exclude_ids)
        qc = None if qcols is None else np.ascontiguousarray(qcols, np.uint32)
        _check(lib().sg_family_batch_aligned(self.h, qmasks, None if qc is None else qc.ctypes.data_as(C.c_void_p), qoff, nq, ex,
                                             C.byref(fp), stride, ids, sc, n))
        return ids, sc, n

    def align(self, qmasks, qoff, fam_ids, fam_off, ap=None):
        """aligner stage with given families: (out_cols, out_masks, results[nq] structured array)."""
        ap = ap or AlignParams()
        qmasks, qoff = np.ascontiguousarray(qmasks, np.uint8), np.ascontiguousarray(qoff, np.uint64)
        fam_ids, fam_off = np.ascontiguousarray(fam_ids, np.uint32), np.ascontiguousarray(fam_off, np.uint64)
        nq = len(qoff) - 1
        oc, om = np.zeros(max(1, len(qmasks)), np.uint32), np.zeros(max(1, len(qmasks)), np.uint8)
        res = np.zeros(nq, RESULT_DTYPE)
        _check(lib().sg_align_batch(self.h, qmasks, qoff, nq, fam_ids if len(fam_ids) else np.zeros(1, np.uint32),
                                    fam_off, C.byref(ap), oc, om, res.ctypes.data_as(C.c_void_p)))
        return oc, om, res

    def run(self, qmasks, qoff, fp=None, ap=None, exclude_ids=None, out=None):
        """famfinder + aligner for a batch, host buffers in and out. `out` = (cols u32[len(qmasks)], masks
        u8[len(qmasks)], results[nq]) lets a caller reuse its output buffers from call to call, as a C++ host does."""
        fp, ap = fp or FamParams(), ap or AlignParams()
        qmasks, qoff = np.ascontiguousarray(qmasks, np.uint8), np.ascontiguousarray(qoff, np.uint64)
        nq = len(qoff) - 1
        if out is not None:
            oc, om, res = out
            assert oc.dtype == np.uint32 and om.dtype == np.uint8 and res.dtype == RESULT_DTYPE
            assert len(oc) >= len(qmasks) and len(om) >= len(qmasks) and len(res) >= nq
            assert oc.flags.c_contiguous and om.flags.c_contiguous and res.flags.c_contiguous
        else:
            oc, om = np.zeros(max(1, len(qmasks)), np.uint32), np.zeros(max(1, len(qmasks)), np.uint8)
            res = np.zeros(nq, RESULT_DTYPE)
        keep, ex = _excl_ptr(exclude_ids)
        _check(lib().sg_run_batch(self.h, qmasks, qoff, nq, ex, C.byref(fp), C.byref(ap), oc, om,
                                  res.ctypes.data_as(C.c_void_p)))
        return oc, om, res


class Session:
    """A batch resident in HBM (sg_session): upload once, run stages, download."""

    def __init__(self, index, max_queries, max_bases):
        self.index = index
        self.h = C.c_void_p()
        _check(lib().sg_session_create(index.h, max_queries, max_bases, C.byref(self.h)))
        self.nq = 0
        self.total = 0

    def close(self):
        if self.h:
            lib().sg_session_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, qmasks, qoff, exclude_ids=None):
        qmasks, qoff = np.ascontiguousarray(qmasks, np.uint8), np.ascontiguousarray(qoff, np.uint64)
        self.nq, self.total = len(qoff) - 1, int(qoff[-1] - qoff[0])
        keep, ex = _excl_ptr(exclude_ids)
        _check(lib().sg_session_upload(self.h, qmasks, qoff, self.nq, ex))

    def find(self, max_results):
        _check(lib().sg_session_find(self.h, max_results))
        self.find_max = min(max_results, self.index.N)

    def turn(self, mode="all"):
        """orientation check on the resident batch; leaves the queries in the chosen orientation"""
        out = np.zeros(self.nq, np.int32)
        _check(lib().sg_session_turn(self.h, TURN_MODES[mode], out.ctypes.data_as(C.c_void_p)))
        return out

    def family(self, fp=None):
        self.fp = fp or FamParams()
        _check(lib().sg_session_family(self.h, C.byref(self.fp)))

    def set_family(self, fam_ids, fam_off):
        _check(lib().sg_session_set_family(self.h, np.ascontiguousarray(fam_ids, np.uint32),
                                           np.ascontiguousarray(fam_off, np.uint64)))

    def align(self, ap=None):
        ap = ap or AlignParams()
        _check(lib().sg_session_align(self.h, C.byref(ap)))

    def run(self, fp=None, ap=None):
        """family finding + alignment in one call (sg_session_run)"""
        self.fp = fp or FamParams()
        ap = ap or AlignParams()
        _check(lib().sg_session_run(self.h, C.byref(self.fp), C.byref(ap)))

    def sync(self):
        _check(lib().sg_session_sync(self.h))

    def download_find(self):
        sc = np.zeros((self.nq, self.find_max), np.int16)
        ids = np.zeros((self.nq, self.find_max), np.uint32)
        nres = np.zeros(self.nq, np.uint32)
        _check(lib().sg_session_download_find(self.h, sc, ids, nres))
        return sc, ids, nres

    def download_family(self):
        stride = self.fp.fs_max + self.fp.fs_req_full + 1
        ids, sc = np.zeros((self.nq, stride), np.uint32), np.zeros((self.nq, stride), np.float32)
        n = np.zeros(self.nq, np.int32)
        _check(lib().sg_session_download_family(self.h, stride, ids, sc, n))
        return ids, sc, n

    def download_align(self):
        oc, om = np.zeros(max(1, self.total), np.uint32), np.zeros(max(1, self.total), np.uint8)
        res = np.zeros(self.nq, RESULT_DTYPE)
        _check(lib().sg_session_download_align(self.h, oc, om, res.ctypes.data_as(C.c_void_p)))
        return oc, om, res

    def stats(self, reset=False):
        st = StageStats()
        _check(lib().sg_session_stats(self.h, C.byref(st), int(reset)))
        return {f: getattr(st, f) for f, _ in StageStats._fields_}

    def timer_start(self):
        """CUDA event on the session's stream; timer_stop() returns the device milliseconds since."""
        _check(lib().sg_session_timer(self.h, 0, None))

    def timer_stop(self):
        ms = C.c_float()
        _check(lib().sg_session_timer(self.h, 1, C.byref(ms)))
        return float(ms.value)

    def dump_graph(self, q, cap_nodes=1 << 17, cap_edges=1 << 19):
        V, E = C.c_uint32(), C.c_uint32()
        col, mask, w = np.zeros(cap_nodes, np.uint32), np.zeros(cap_nodes, np.uint8), np.zeros(cap_nodes, np.float32)
        po, pr = np.zeros(cap_nodes + 1, np.uint32), np.zeros(cap_edges, np.uint32)
        _check(lib().sg_session_dump_graph(self.h, q, cap_nodes, cap_edges, C.byref(V), C.byref(E), col, mask, w, po,
                                           pr))
        return dict(V=V.value, E=E.value, col=col[:V.value], mask=mask[:V.value], weight=w[:V.value],
                    pred_off=po[:V.value + 1], preds=pr[:E.value])
