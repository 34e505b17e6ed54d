"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/sina_b200.h declares; no compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import pytest

import sina_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sina_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z_0-9]+)\s*\(", src)))


def test_header_matches_binding_list():
    assert header_symbols() == sorted(sina_b200.EXPORTS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(sina_b200.LIB_PATH):
        pytest.fail("libsina_b200.so not built: run __graft_entry__.build()")
    L = C.CDLL(sina_b200.LIB_PATH)
    for name in header_symbols():
        assert hasattr(L, name), name


def test_defaults_match_reference_options():
    """sg_default_*_params carry the reference's option defaults (famfinder.cpp:155-195, align.cpp:232-259)."""
    L = sina_b200.lib()
    fp, ap = sina_b200.FamParams(0, 0, 0, 0, 0, 0, 0, 0, 0, 1), sina_b200.AlignParams(0, 0, 0, 0, 0, 1, 1, 1, 1)
    L.sg_default_fam_params(C.byref(fp))
    L.sg_default_align_params(C.byref(ap))
    assert (fp.fs_min, fp.fs_max, fp.fs_min_len, fp.fs_req_full, fp.fs_full_len, fp.fs_req_gaps, fp.fs_req,
            fp.leave_query_out) == (40, 40, 150, 1, 1400, 10, 1, 0)
    assert abs(fp.fs_msc - 0.7) < 1e-6 and fp.fs_msc_max == 2.0
    assert (ap.match_score, ap.mismatch_score, ap.gap_penalty, ap.gap_ext_penalty, ap.fs_weight) == (2, -1, 5, 2, 1)
    assert (ap.overhang, ap.lowercase, ap.insertion, ap.realign) == (0, 0, 0, 0)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail loudly, never fall back."""
    if sina_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    import numpy as np
    with pytest.raises(sina_b200.SinaB200Error):
        sina_b200.Index(np.array([1, 2, 4, 8], np.uint8), np.arange(4, dtype=np.uint32), np.array([0, 4], np.uint64),
                        10, k=2)


def test_product_never_imports_oracle():
    """the product package must not reference oracle/ (checked textually over sina_b200/)."""
    for dp, _, files in os.walk(os.path.join(ROOT, "sina_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
