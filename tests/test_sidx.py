"""The .sidx index cache: sina_b200/host/sidx.cpp against the reference's own file code -- vlimap::write /
read / invert of src/idset.h compiled in place, driven like kmer_search::impl::store / try_load (src/kmer_search.cpp:278-351).
CPU only: the lists come from the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O
from sina_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "sina_b200", "libsina_host.so")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(HOST_SO):
        pytest.skip("libsina_host.so not built")
    L = C.CDLL(HOST_SO)
    L.sina_sidx_write.restype = C.c_int
    L.sina_sidx_write.argtypes = [C.c_char_p, C.c_uint, C.c_int, C.c_void_p, C.c_uint32, u64p, u32p, C.c_uint64]
    L.sina_sidx_read.restype = C.c_int64
    L.sina_sidx_read.argtypes = [C.c_char_p] + [C.c_void_p] * 8 + [C.c_uint64]
    return L


def write_ours(host, path, k, nofast, names, off, ids):
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    ids = np.ascontiguousarray(ids if len(ids) else np.zeros(1), np.uint32)
    assert host.sina_sidx_write(str(path).encode(), k, int(nofast), C.cast(arr, C.c_void_p), len(names),
                                np.ascontiguousarray(off, np.uint64), ids, len(off) - 1) == 0


def read_ours(host, path):
    k, nf, n, nk = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    total = host.sina_sidx_read(str(path).encode(), C.byref(k), C.byref(nf), C.byref(n), C.byref(nk), None, None, None, None, 0)
    assert total >= 0
    kmers = np.zeros(max(1, nk.value), np.uint32)
    off = np.zeros(nk.value + 1, np.uint64)
    ids = np.zeros(max(1, total), np.uint32)
    names = C.create_string_buffer(64 * (n.value + 1))
    host.sina_sidx_read(str(path).encode(), None, None, None, None, kmers.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                        ids.ctypes.data_as(C.c_void_p), names, len(names))
    return dict(k=k.value, nofast=nf.value, n=n.value, kmers=kmers[:nk.value], off=off, ids=ids[:total],
                names=names.value.decode().split("\n")[:-1])


def same_file(a, b):
    """byte for byte, except the compiler's padding inside idx_header (bytes 10-11 and 18-23)"""
    x, y = bytearray(open(a, "rb").read()), bytearray(open(b, "rb").read())
    for r in (slice(10, 12), slice(18, 24)):
        x[r] = y[r] = b"\0" * (r.stop - r.start)
    return x == y


@pytest.mark.parametrize("k,nofast", [(4, False), (6, False), (3, True)])
def test_sidx_file_equals_reference(host, orc, ref, tmp_path, k, nofast):
    """same index, same bytes as kmer_search::impl::store would write (k = 3 / 4: most lists are inverted); the
    reference's reader loads our file and ranks with it like with the index it built itself; our reader gets the
    lists back from the reference's file"""
    tree, m, c, o = synth.synth_msa(300, W=900, L=260, seed=5)
    msa = O.MSA(m, c, o, 900)
    names = ["ref%d" % i for i in range(msa.N)]
    oix = orc.index_build(msa, k, nofast)
    off, ids = orc.index_lists(oix)
    db = ref.db(msa)
    rix = ref.kidx_build(db, k, nofast)
    ref.kidx_store(rix, names, tmp_path / "ref.sidx")
    write_ours(host, tmp_path / "ours.sidx", k, nofast, names, off, ids)
    assert same_file(tmp_path / "ours.sidx", tmp_path / "ref.sidx")
    assert sum(1 for v in range(len(off) - 1) if off[v + 1] - off[v] > msa.N // 2) > 0 or k > 4   # inverted lists exercised
    # reference reader on our file
    lix = ref.kidx_load(db, tmp_path / "ours.sidx", k, nofast)
    assert lix is not None
    qm, qo = synth.synth_queries(tree, 6, "full", seed=3)
    for i in range(6):
        q = O.decode(qm[int(qo[i]):int(qo[i + 1])])
        s1, i1, _ = ref.find(rix, q, 25)
        s2, i2, _ = ref.find(lix, q, 25)
        assert (s1 == s2).all() and (i1 == i2).all()
    ref.kidx_free(lix)
    # our reader on the reference's file
    got = read_ours(host, tmp_path / "ref.sidx")
    assert (got["k"], got["nofast"], got["n"], got["names"]) == (k, int(nofast), msa.N, names)
    want_kmers = [v for v in range(len(off) - 1) if off[v + 1] > off[v]]
    assert list(got["kmers"]) == want_kmers
    for j, v in enumerate(want_kmers):
        assert (got["ids"][int(got["off"][j]):int(got["off"][j + 1])] == ids[int(off[v]):int(off[v + 1])]).all(), v
    ref.kidx_free(rix)
    ref.db_free(db)
    orc.index_free(oix)


def test_sidx_vlimap_sizes_and_fills(host, ref, tmp_path):
    """the size / fill grid of the reference's own posting-list test (src/unit_tests/idset_test.cpp:78-80: sizes 0, 255,
    256, 257, 10000; fill 0, 10, 50, 100 %; three seeds): every list written by us equals the reference's vlimap::write
    of the same ids, and read back through the reference's vlimap::read + increment it marks exactly those ids"""
    for N in (0, 255, 256, 257, 10000):
        kmers, off, ids = [], [0], []
        j = 0
        for fill in (0, 10, 50, 100):
            for seed in (132456, 54321, 242424):
                rng = np.random.default_rng(seed + N + fill)
                sel = np.nonzero(rng.random(N) * 100 < fill)[0].astype(np.uint32)
                kmers.append(j)
                ids.append(sel)
                off.append(off[-1] + len(sel))
                j += 1
        k = 2   # 16 k-mer slots hold the 12 lists
        flat = np.concatenate(ids) if N else np.zeros(0, np.uint32)
        names = ["s%d" % i for i in range(N)]
        rix = ref.kidx_from_lists(N, k, True, kmers, off, flat)
        ref.kidx_store(rix, names, tmp_path / "ref.sidx")
        full_off = np.array(off + [off[-1]] * (16 - len(kmers)), np.uint64)
        write_ours(host, tmp_path / "ours.sidx", k, True, names, full_off, flat)
        assert same_file(tmp_path / "ours.sidx", tmp_path / "ref.sidx"), N
        lix = ref.kidx_load(None, tmp_path / "ours.sidx", k, True)
        assert lix is not None
        for j, sel in enumerate(ids):
            if len(sel) == 0:
                continue   # empty lists are not stored (kmer_search.cpp:291-295)
            sc, _ = ref.kidx_list_scores(lix, j, N)
            want = np.zeros(N, np.int16)
            want[sel] = 1
            assert (sc == want).all(), (N, j)
        ref.kidx_free(lix)
        ref.kidx_free(rix)
