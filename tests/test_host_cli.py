"""GPU: the `sina` command line and the per-tray stage functors of sina_b200/host against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from sina_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sina_b200", "bin")
FAM = dict(fs_min=20, fs_max=20, fs_min_len=100, fs_full_len=240, fs_req_gaps=5)
FAM_ARGS = ["--fs-kmer-len", "6", "--fs-min", "20", "--fs-max", "20", "--fs-min-len", "100", "--fs-full-len", "240", "--fs-req-gaps", "5"]


def write_fasta(path, names, seqs, width=0):
    with open(path, "w") as f:
        for n, s in zip(names, seqs):
            f.write(">%s\n" % n)
            if width:
                for i in range(0, len(s), width):
                    f.write(s[i:i + width] + "\n")
            else:
                f.write(s + "\n")


def read_fasta(path):
    out, name = {}, None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:].split()[0]
            out[name] = ""
        elif name:
            out[name] += line
    return out


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    tree, m, c, o = synth.synth_msa(300, W=900, L=260, seed=5)
    msa = O.MSA(m, c, o, 900)
    nq = 37
    qm, qo = synth.synth_queries(tree, nq, "full", seed=23)
    write_fasta(d / "ref.fasta", ["ref%d" % i for i in range(msa.N)], [msa.row_string(i) for i in range(msa.N)], width=70)
    qs = [O.decode(qm[int(qo[i]):int(qo[i + 1])]) for i in range(nq)]
    qs[3] = qs[3].lower()          # case is ignored unless --lowercase original
    qs[5] = qs[5][:40] + "n" + qs[5][41:]
    write_fasta(d / "q.fasta", ["q%d" % i for i in range(nq)], qs, width=60)
    qmasks = [O.encode(s) for s in qs]
    return d, msa, qmasks


def oracle_strings(orc, msa, qmasks, ap_kw, fp_kw=FAM, k=6):
    oix = orc.index_build(msa, k, 0)
    qoff = np.zeros(len(qmasks) + 1, np.uint64)
    qoff[1:] = np.cumsum([len(q) for q in qmasks])
    qm = np.concatenate(qmasks)
    res, oc, om, cells, posts, nt = orc.run_batch(oix, msa, qm, qoff, O.FamParams(**fp_kw), O.AlignParams(**ap_kw))
    out = []
    for i in range(len(qmasks)):
        a, n = int(qoff[i]), res[i].n_out
        out.append(O.render(om[a:a + n], oc[a:a + n], msa.W) if res[i].status in (0, 1) else None)
    orc.index_free(oix)
    return out, res


@pytest.mark.parametrize("cli,ap_kw", [
    ([], {}),
    (["--overhang", "edge", "--lowercase", "unaligned", "--batch-size", "8"], dict(overhang=2, lowercase=2)),
    (["--lowercase", "original", "--pen-gap", "4.3", "--pen-gapext", "1.1", "--match-score", "1.7", "--mismatch-score", "-0.9",
      "--fs-weight", "0.5", "--line-length", "100"],
     dict(lowercase=1, gap_penalty=4.3, gap_ext_penalty=1.1, match_score=1.7, mismatch_score=-0.9, fs_weight=0.5)),
    (["--insertion", "forbid", "--overhang", "remove"], dict(insertion=1, overhang=1)),
])
def test_cli_matches_oracle(orc, data, cli, ap_kw):
    d, msa, qmasks = data
    out = d / "out.fasta"
    r = subprocess.run([os.path.join(BIN, "sina"), "-i", str(d / "q.fasta"), "-o", str(out), "--db", str(d / "ref.fasta"),
                        "--fs-engine", "internal", "--gpus", "1"] + FAM_ARGS + cli, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "sequences/s" in r.stderr
    got = read_fasta(out)
    want, res = oracle_strings(orc, msa, qmasks, ap_kw)
    assert list(got) == ["q%d" % i for i in range(len(qmasks)) if want[i] is not None]  # input order kept
    for i, w in enumerate(want):
        if w is not None:
            assert got["q%d" % i] == w, i


def test_stage_functors_per_tray(orc, data):
    """famfinder::operator()(tray) / aligner::operator()(tray) / kmer_search::find one query at a time"""
    d, msa, qmasks = data
    r = subprocess.run([os.path.join(BIN, "stage_dump"), str(d / "ref.fasta"), str(d / "q.fasta")] + FAM_ARGS,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().split("\n")
    assert lines[0] == "size %d" % msa.N
    want, res = oracle_strings(orc, msa, qmasks, {})
    oix = orc.index_build(msa, 6, 0)
    blocks = [lines[i:i + 6] for i in range(1, len(lines), 6)]
    assert len(blocks) == len(qmasks)
    for i, b in enumerate(blocks):
        assert b[0] == "query q%d" % i
        sc, ids, _ = orc.find(oix, qmasks[i], 5)
        assert b[1] == "find" + "".join(" ref%d:%d" % (j, s) for j, s in zip(ids, sc))
        n, fids, fsc = orc.family(oix, msa, qmasks[i], O.FamParams(**FAM))
        assert b[2] == "family" + "".join(" ref%d:%d" % (j, s) for j, s in zip(fids, fsc))
        assert b[3] == "aligned " + want[i]
        assert b[4] == "attrs %d %d %d" % (res[i].qual, res[i].head, res[i].tail)
        assert b[5].startswith("log scoring: raw=")
    orc.index_free(oix)


def test_cli_soft_failures_and_copy(orc, data, tmp_path):
    """too few relatives => not written (src/famfinder.cpp:486-491); query contained in a reference => alignment copied
    (src/align.cpp:349-388); with --realign that reference is dropped instead"""
    d, msa, qmasks = data
    m5, c5 = msa.row(5)
    sub = O.decode(m5[10:200])
    write_fasta(tmp_path / "q.fasta", ["contained", "tiny", "normal"], [sub, "A", O.decode(qmasks[0])])
    out = tmp_path / "out.fasta"
    base = [os.path.join(BIN, "sina"), "-i", str(tmp_path / "q.fasta"), "-o", str(out), "--db", str(d / "ref.fasta")] + FAM_ARGS
    r = subprocess.run(base + ["--show-log"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = read_fasta(out)
    assert list(got) == ["contained", "normal"]
    assert "too few relatives" in r.stderr and "1 sequences were not aligned" in r.stderr
    w1, res1 = oracle_strings(orc, msa, [O.encode(sub)], {})
    assert res1[0].status == 1 and got["contained"] == w1[0]
    assert "copied alignment" in r.stderr
    r = subprocess.run(base + ["--realign"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0
    w2, res2 = oracle_strings(orc, msa, [O.encode(sub)], dict(realign=1))
    assert res2[0].status == 0 and read_fasta(out)["contained"] == w2[0]


def test_cli_turn(orc, data, tmp_path):
    """--turn all / revcomp (src/famfinder.cpp:311-378): queries given in any of the four orientations come out aligned
    like their forward form; without --turn the reference-side semantics keep them as they are"""
    d, msa, qmasks = data
    comp = str.maketrans("ACGUacgu", "UGCAugca")
    fwd = [O.decode(q) for q in qmasks[:12]]
    given = []
    for i, s in enumerate(fwd):
        t = i % 4
        s2 = s[::-1] if t & 1 else s
        given.append(s2.translate(comp) if t & 2 else s2)
    write_fasta(tmp_path / "q.fasta", ["q%d" % i for i in range(12)], given)
    out = tmp_path / "out.fasta"
    base = [os.path.join(BIN, "sina"), "-i", str(tmp_path / "q.fasta"), "-o", str(out), "--db", str(d / "ref.fasta")] + FAM_ARGS
    r = subprocess.run(base + ["--turn", "all"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = read_fasta(out)
    oix = orc.index_build(msa, 6, 0)
    turns = [orc.turn_check(oix, O.encode(g), True)[0] for g in given]
    orc.index_free(oix)

    def apply(s, t):
        s2 = s[::-1] if t & 1 else s
        return s2.translate(comp) if t & 2 else s2
    want, res = oracle_strings(orc, msa, [O.encode(apply(g, t)) for g, t in zip(given, turns)], {})
    for i, w in enumerate(want):
        if w is not None:
            assert got["q%d" % i] == w, i
    r = subprocess.run(base + ["--turn", "revcomp"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = read_fasta(out)
    oix = orc.index_build(msa, 6, 0)
    turns = [orc.turn_check(oix, O.encode(g), False)[0] for g in given]
    orc.index_free(oix)
    want, res = oracle_strings(orc, msa, [O.encode(apply(g, t)) for g, t in zip(given, turns)], {})
    for i, w in enumerate(want):
        if w is not None:
            assert got["q%d" % i] == w, i


def test_align_test_mimic_12x12_realign(orc, tmp_path):
    """BASELINE configs[0], the reference's tests/align.test workload: every 1000th sequence of a database (12 of them)
    is extracted and re-aligned against the extract itself with --realign (each query is contained in the database, so
    its own row is dropped from the family, src/align.cpp:337-348); the test asserts 'align 12 sequences', exit 0 and that
    writing to stdout gives the same bytes as writing to a file (tests/align.test:11-32). The ARB fixture cannot be read
    here, so the database is synthetic; unlike the reference's test the aligned strings are pinned too."""
    tree, m, c, o = synth.synth_msa(12000, W=5000, L=1500, seed=77)
    big = O.MSA(m, c, o, 5000)
    pick = list(range(0, 12000, 1000))
    rows = [big.row_string(i) for i in pick]
    names = ["seq%d" % i for i in pick]
    write_fasta(tmp_path / "extracted.fasta", names, rows, width=80)
    msa = O.MSA.from_rows(rows)
    base = [os.path.join(BIN, "sina"), "-i", str(tmp_path / "extracted.fasta"), "--preserve-order", "--realign", "--db",
            str(tmp_path / "extracted.fasta"), "--fs-engine", "internal", "--gpus", "1"]
    r1 = subprocess.run(base + ["-o", str(tmp_path / "aligned.fasta")], capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stderr
    assert "align 12 sequences" in r1.stderr
    r2 = subprocess.run(base, capture_output=True, timeout=600)          # FASTA -> stdout
    assert r2.returncode == 0 and b"align 12 sequences" in r2.stderr
    assert r2.stdout == open(tmp_path / "aligned.fasta", "rb").read()     # cmp aligned.fasta aligned.2.fasta
    got = read_fasta(tmp_path / "aligned.fasta")
    qmasks = [msa.row(i)[0] & 15 for i in range(12)]
    want, res = oracle_strings(orc, msa, qmasks, dict(realign=1), fp_kw={}, k=10)
    assert list(got) == [names[i] for i in range(12) if want[i] is not None]
    for i in range(12):
        if want[i] is not None:
            assert res[i].status == 0 and got[names[i]] == want[i], i


def test_cli_filter_weights(orc, data, tmp_path):
    """--filter-weights FILE: positional column weights -> scoring_scheme_weighted (src/align.cpp:409-415)"""
    d, msa, qmasks = data
    rng = np.random.default_rng(5)
    w = (0.5 - np.log(rng.uniform(1e-4, 0.95, msa.W))).astype(np.float32)
    with open(tmp_path / "weights.txt", "w") as f:
        f.write("\n".join("%.9g" % x for x in w) + "\n")
    out = tmp_path / "out.fasta"
    r = subprocess.run([os.path.join(BIN, "sina"), "-i", str(d / "q.fasta"), "-o", str(out), "--db", str(d / "ref.fasta"),
                        "--filter-weights", str(tmp_path / "weights.txt")] + FAM_ARGS, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    orc.set_column_weights(w)
    try:
        want, res = oracle_strings(orc, msa, qmasks, {})
    finally:
        orc.set_column_weights(None)
    got = read_fasta(out)
    for i, s in enumerate(want):
        if s is not None:
            assert got["q%d" % i] == s, i
    plain, _ = oracle_strings(orc, msa, qmasks, {})
    assert any(a != b for a, b in zip(want, plain))   # the weights do change alignments
    r = subprocess.run([os.path.join(BIN, "sina"), "-i", str(d / "q.fasta"), "-o", str(out), "--db", str(d / "ref.fasta"),
                        "--filter", "pos_var"] + FAM_ARGS, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "filter-weights" in r.stderr


def test_cli_sidx_cache_and_index_order(orc, data, tmp_path):
    """like kmer_search::impl::impl (src/kmer_search.cpp:213-242) the command line leaves `<db>.sidx` next to the database:
    the file equals the one the reference's own code writes for the same index; and a .sidx found next to a database
    defines the index ORDER (ids = positions in its name list, :289-291), which decides ties between equal k-mer scores"""
    if not O.have_ref():
        pytest.skip("compiled reference (oracle/_ref) not available")
    d, msa, qmasks = data
    import shutil
    db = tmp_path / "ref.fasta"
    shutil.copy(d / "ref.fasta", db)
    names = ["ref%d" % i for i in range(msa.N)]
    base = [os.path.join(BIN, "sina"), "-i", str(d / "q.fasta"), "-o", str(tmp_path / "out.fasta"), "--db", str(db)] + FAM_ARGS
    r = subprocess.run(base, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(str(db) + ".sidx")
    ref = O.Ref()
    rdb = ref.db(msa)
    rix = ref.kidx_build(rdb, 6, 0)
    ref.kidx_store(rix, names, tmp_path / "want.sidx")
    x, y = bytearray(open(str(db) + ".sidx", "rb").read()), bytearray(open(tmp_path / "want.sidx", "rb").read())
    for sl in (slice(10, 12), slice(18, 24)):   # padding inside idx_header
        x[sl] = y[sl] = b"\0" * (sl.stop - sl.start)
    assert x == y
    ref.kidx_free(rix)
    ref.db_free(rdb)
    first = read_fasta(tmp_path / "out.fasta")
    # a cache that lists the sequences in another order: ids follow it
    perm = np.random.default_rng(8).permutation(msa.N)
    rows = [msa.row_string(int(i)) for i in perm]
    pmsa = O.MSA.from_rows(rows)
    pdb = ref.db(pmsa)
    pix = ref.kidx_build(pdb, 6, 0)
    ref.kidx_store(pix, [names[int(i)] for i in perm], str(db) + ".sidx")
    ref.kidx_free(pix)
    ref.db_free(pdb)
    r = subprocess.run(base, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = read_fasta(tmp_path / "out.fasta")
    want, res = oracle_strings(orc, pmsa, qmasks, {})
    for i, w in enumerate(want):
        if w is not None:
            assert got["q%d" % i] == w, i
    assert first.keys() == got.keys()


def test_cli_search_stage(data, tmp_path):
    """--search: search_filter after the aligner (src/sina.cpp:519-527, src/search_filter.cpp:244-330); `nearest_slv` lists
    the max_result nearest references by identity (name~score, score with three decimals as in the reference)"""
    if not O.have_ref():
        pytest.skip("compiled reference (oracle/_ref) not available")
    d, msa, qmasks = data
    out = tmp_path / "out.fasta"
    r = subprocess.run([os.path.join(BIN, "sina"), "-i", str(d / "q.fasta"), "-o", str(out), "--db", str(d / "ref.fasta"),
                        "--search", "--search-kmer-len", "6", "--search-max-result", "5", "--search-min-sim", "0.5",
                        "--search-kmer-candidates", "50", "--meta-fmt", "comment"] + FAM_ARGS,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    # the aligned sequences and their nearest_slv comment lines
    recs, name = {}, None
    for line in open(out):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:].split()[0]
            recs[name] = {"seq": "", "nearest": None}
        elif line.startswith(";"):
            k, _, v = line[1:].strip().partition("=")
            if k == "nearest_slv":
                recs[name]["nearest"] = v
        elif name:
            recs[name]["seq"] += line
    ref = O.Ref()
    names = ["ref%d" % i for i in range(msa.N)]
    rdb = ref.db(O.MSA(msa.masks, msa.cols, msa.off, msa.W, names=names))
    rix = ref.kidx_build(rdb, 6, 0)
    checked = 0
    for qn, rec in recs.items():
        a = O._CHAR2MASK[np.frombuffer(rec["seq"].encode(), np.uint8)]
        cols = np.nonzero(a > 0)[0].astype(np.uint32)
        masks = a[cols].astype(np.uint8)
        ids, sc = ref.search(rix, masks, cols, 50, 5, 0.5)
        want = "".join("%s~%.3f " % (names[i], s) for i, s in zip(ids, sc))
        assert (rec["nearest"] or "") == want.strip() or (rec["nearest"] or "") == want, (qn, rec["nearest"], want)
        checked += len(ids) > 0
    assert checked >= 30
    ref.kidx_free(rix)
    ref.db_free(rdb)


def test_cli_calc_idty(data, tmp_path):
    """--calc-idty (src/align.cpp:443-453): align_ident_slv = 100 x the highest identity (optimistic, relative to the
    overlap) of the aligned sequence with any relative of its family; 100 for a copied alignment"""
    if not O.have_ref():
        pytest.skip("compiled reference (oracle/_ref) not available")
    d, msa, qmasks = data
    out = tmp_path / "out.fasta"
    r = subprocess.run([os.path.join(BIN, "sina"), "-i", str(d / "q.fasta"), "-o", str(out), "--db", str(d / "ref.fasta"),
                        "--calc-idty", "--meta-fmt", "comment"] + FAM_ARGS, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    recs, name = {}, None
    for line in open(out):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:].split()[0]
            recs[name] = {"seq": "", "idty": None}
        elif line.startswith(";"):
            k, _, v = line[1:].strip().partition("=")
            if k == "align_ident_slv":
                recs[name]["idty"] = v
        elif name:
            recs[name]["seq"] += line
    ref = O.Ref()
    rdb = ref.db(msa)
    rix = ref.kidx_build(rdb, 6, 0)
    checked = 0
    for i, qm in enumerate(qmasks):
        rec = recs.get("q%d" % i)
        if rec is None:
            continue
        n, ids, _ = ref.family(rix, O.decode(qm), O.FamParams(**FAM))
        a = O._CHAR2MASK[np.frombuffer(rec["seq"].encode(), np.uint8)]
        cols = np.nonzero(a > 0)[0].astype(np.uint32)
        masks = a[cols].astype(np.uint8)
        best = np.float32(0)
        for rid in ids:
            bm, bc = msa.row(int(rid))
            v = np.float32(ref.compare(masks, cols, bm, bc, 0, 0, 3, False))
            if v > best:
                best = v
        want = "%.9g" % float(np.float32(100) * best)
        assert rec["idty"] == want, (i, rec["idty"], want)
        checked += 1
    assert checked >= 30
    ref.kidx_free(rix)
    ref.db_free(rdb)
    # --min-idty (src/rw_fasta.cpp:405-414): sequences below the threshold are not written
    vals = sorted(float(rec["idty"]) for rec in recs.values())
    thr = vals[len(vals) // 2]
    out2 = tmp_path / "out2.fasta"
    r = subprocess.run([os.path.join(BIN, "sina"), "-i", str(d / "q.fasta"), "-o", str(out2), "--db", str(d / "ref.fasta"),
                        "--calc-idty", "--min-idty", "%.9g" % thr] + FAM_ARGS, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    got = set(read_fasta(out2))
    assert got == {n for n, rec in recs.items() if not (np.float32(thr) > np.float32(float(rec["idty"])))}
    assert 0 < len(got) < len(recs)
