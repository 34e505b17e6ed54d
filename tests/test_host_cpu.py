"""Host-side mirror of the reference's stage interfaces (sina_b200/host): unit checks with the reference's own
cseq vectors, the option surface, and the loud failure without a GPU. No device needed."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sina_b200", "bin")


def have_gpu():
    import sina_b200
    try:
        return sina_b200.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(os.path.join(BIN, "sina")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "sina_b200", "host"), "-j", "4"], check=True, capture_output=True)


def run(args, **kw):
    return subprocess.run([os.path.join(BIN, args[0])] + args[1:], capture_output=True, text=True, timeout=120, **kw)


def test_host_unit(tmp_path):
    r = run(["host_unit", str(tmp_path)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok ")


def test_cli_help_and_version():
    r = run(["sina", "--help"])
    assert r.returncode == 0
    for name in ("--db", "--fs-engine", "--fs-kmer-len", "--fs-min", "--fs-max", "--fs-msc", "--fs-req", "-i [ --in ]", "-o [ --out ]"):
        assert name in r.stderr, name
    r = run(["sina", "--help-all"])
    for name in ("--pen-gap", "--pen-gapext", "--match-score", "--mismatch-score", "--overhang", "--lowercase", "--insertion",
                 "--fs-kmer-no-fast", "--fs-kmer-mm", "--fs-msc-max", "--fs-leave-query-out", "--realign", "--fs-weight",
                 "--preserve-order", "--max-in-flight"):
        assert name in r.stderr, name
    assert run(["sina", "--version"]).returncode == 0


@pytest.mark.parametrize("args,msg", [
    (["--fs-engine", "pt-server", "--db", "x"], "pt-server is not supported"),
    (["--db", "x", "--fs-no-graph"], "not supported"),
    (["--db", "x", "--use-subst-matrix"], "not supported"),
    (["--db", "x", "--filter", "f"], "not supported"),
    (["--db", "x", "--insertion", "sideways"], "insertion type must be one of"),
    (["--db", "x", "--turn", "sideways"], "Turn type must be one of"),
    (["--db", "x", "--search", "--search-all"], "not supported"),
    (["--db", "x", "--search", "--lca-fields", "tax_slv"], "not supported"),
    (["--db", "x", "--search", "--search-cover", "abs", "--search-correction", "jc"], "only fractional identity"),
    (["--db", "x", "--search", "--search-iupac", "sometimes"], "iupac matching must be"),
    (["--db", "x", "--bogus"], "unrecognised option"),
    (["-i", "q"], "Must have reference database"),
])
def test_cli_rejects(args, msg):
    """unsupported reference options are rejected (exit 1, like a reference configuration error), never ignored"""
    r = run(["sina"] + args)
    assert r.returncode == 1
    assert msg in r.stderr, r.stderr


def test_cli_fails_loudly_without_gpu(tmp_path):
    if have_gpu():
        pytest.skip("a CUDA device is present")
    q, d = tmp_path / "q.fa", tmp_path / "r.fa"
    q.write_text(">q\nAGCU\n")
    d.write_text(">r\nAG-CU\n")
    r = run(["sina", "-i", str(q), "--db", str(d)])
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    assert r.stdout == ""  # nothing was "aligned" by some fallback


def test_cli_prealigned_passthrough_keeps_order(tmp_path):
    """--prealigned (no device needed): reader -> render pool -> writer over many small batches keeps the input order
    (the reference's sequencer_node, src/sina.cpp:529-538) and re-emits the aligned rows, wrapped by --line-length"""
    import random
    rnd = random.Random(5)
    names, rows = [], []
    for i in range(57):
        names.append("s%03d" % i)
        rows.append("".join(rnd.choice("ACGU-----") for _ in range(200)))
    with open(tmp_path / "in.fasta", "w") as f:
        for n, r in zip(names, rows):
            f.write(">%s\n%s\n" % (n, r))
    out = tmp_path / "out.fasta"
    r = run(["sina", "--prealigned", "-i", str(tmp_path / "in.fasta"), "-o", str(out), "--batch-size", "3", "--line-length", "70"])
    assert r.returncode == 0, r.stderr
    got_names, got = [], {}
    for line in open(out):
        line = line.rstrip("\n")
        if line.startswith(">"):
            got_names.append(line[1:].split()[0])
            got[got_names[-1]] = []
        else:
            got[got_names[-1]].append(line)
    assert got_names == names
    for n, row in zip(names, rows):
        assert all(len(x) <= 70 for x in got[n])
        assert "".join(got[n]) == row


def test_cli_gzip_in_and_out(tmp_path):
    """file names ending in .gz are gzip streams on both sides (src/rw_fasta.cpp:200-202,358-360). --prealigned needs no
    device: compressed input, compressed output (one gzip member per 256 records, written in input order), and the
    same bytes as the uncompressed run once decompressed"""
    import gzip, random
    rnd = random.Random(9)
    text = ""
    for i in range(700):   # three members per batch of 600, two batches
        text += ">g%04d some description\n%s\n" % (i, "".join(rnd.choice("ACGU------") for _ in range(300)))
    with open(tmp_path / "in.fasta", "w") as f:
        f.write(text)
    with gzip.open(tmp_path / "in.fasta.gz", "wt") as f:
        f.write(text)
    plain, packed = tmp_path / "out.fasta", tmp_path / "out.fasta.gz"
    r = run(["sina", "--prealigned", "-i", str(tmp_path / "in.fasta"), "-o", str(plain), "--batch-size", "600", "--line-length", "80"])
    assert r.returncode == 0, r.stderr
    r = run(["sina", "--prealigned", "-i", str(tmp_path / "in.fasta.gz"), "-o", str(packed), "--batch-size", "600", "--line-length", "80"])
    assert r.returncode == 0, r.stderr
    raw = open(packed, "rb").read()
    assert raw[:2] == b"\x1f\x8b" and len(raw) < os.path.getsize(plain)
    assert gzip.decompress(raw) == open(plain, "rb").read()
    assert raw.count(b"\x1f\x8b\x08") >= 4                      # several members
    # an empty input still gives a valid (empty) stream, and block-wise input refuses compressed files
    open(tmp_path / "empty.fasta", "w").close()
    r = run(["sina", "--prealigned", "-i", str(tmp_path / "empty.fasta"), "-o", str(tmp_path / "e.fasta.gz")])
    assert r.returncode == 0, r.stderr
    assert gzip.decompress(open(tmp_path / "e.fasta.gz", "rb").read()) == b""
    r = run(["sina", "--prealigned", "-i", str(tmp_path / "in.fasta.gz"), "-o", "/dev/null", "--fasta-block", "1000", "--fasta-idx", "1"])
    assert r.returncode != 0 and "compressed" in r.stderr


def test_cli_meta_fmt_csv(tmp_path):
    """--meta-fmt csv (src/rw_fasta.cpp:362-372,484-515): plain FASTA headers, the attributes in <out>.csv with CRLF lines,
    the column names taken from the first record written, values with commas / quotes escaped (:379-392)"""
    with open(tmp_path / "in.fasta", "w") as f:
        f.write(">a desc A\n; turn = none\n; align_quality_slv = 97\nAC--GU\n>b with, comma\n--ACGU\n>c say \"hi\"\n; turn = all\nAC-G-U\n")
    out = tmp_path / "res.fasta"
    r = run(["sina", "--prealigned", "--meta-fmt", "csv", "-i", str(tmp_path / "in.fasta"), "-o", str(out), "--batch-size", "2"])
    assert r.returncode == 0, r.stderr
    assert open(out).read() == ">a desc A\nAC--GU\n>b with, comma\n--ACGU\n>c say \"hi\"\nAC-G-U\n"
    assert open(tmp_path / "res.csv", "rb").read() == (b"name,align_quality_slv,full_name,turn\r\n"
                                                       b"a,97,desc A,none\r\n"
                                                       b"b,\"with, comma\"\r\n"
                                                       b"c,\"say \"\"hi\"\"\",all\r\n")


def test_cli_mmap_output_matches_pwrite(tmp_path):
    """SINA_B200_MMAP_OUT=1: the output pool fills mappings of the reserved byte ranges instead of calling pwrite; same
    bytes, over slices that start and end anywhere inside a page, and an empty result is an empty file"""
    import random
    rnd = random.Random(11)
    with open(tmp_path / "in.fasta", "w") as f:
        for i in range(1500):
            f.write(">m%04d d\n%s\n" % (i, "".join(rnd.choice("ACGU-------") for _ in range(rnd.randint(50, 6000)))))
    outs = []
    for mode in ("0", "1"):
        out = tmp_path / ("out%s.fasta" % mode)
        r = run(["sina", "--prealigned", "-i", str(tmp_path / "in.fasta"), "-o", str(out), "--batch-size", "700", "--line-length", "77"],
                env=dict(os.environ, SINA_B200_MMAP_OUT=mode))
        assert r.returncode == 0, r.stderr
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1] and len(outs[0]) > 1500 * 50
    open(tmp_path / "empty.fasta", "w").close()
    r = run(["sina", "--prealigned", "-i", str(tmp_path / "empty.fasta"), "-o", str(tmp_path / "e.fasta")], env=dict(os.environ, SINA_B200_MMAP_OUT="1"))
    assert r.returncode == 0 and os.path.getsize(tmp_path / "e.fasta") == 0


def test_cli_fasta_block_and_idx(tmp_path):
    """--fasta-block B --fasta-idx i (src/rw_fasta.cpp:209-216,237-242): seek to byte B*i, skip to the next title line, read
    records until the previous one ended past byte B*(i+1). No GPU needed: --prealigned passes the sequences through."""
    rng = np.random.default_rng(3)
    names, seqs, starts, text = [], [], [], ""
    for i in range(40):
        n = int(rng.integers(30, 400))
        s = "".join(rng.choice(list("ACGU-"), size=n))
        names.append("s%d" % i)
        starts.append(len(text))
        text += ">s%d\n" % i
        for j in range(0, n, 60):
            text += s[j:j + 60] + "\n"
    (tmp_path / "in.fasta").write_text(text)
    B = 700
    seen = []
    for idx in range(len(text) // B + 1):
        out = tmp_path / ("out%d.fasta" % idx)
        r = run(["sina", "-i", str(tmp_path / "in.fasta"), "-o", str(out), "--prealigned", "--fasta-block", str(B), "--fasta-idx", str(idx)])
        assert r.returncode == 0, r.stderr
        got = [l[1:].split()[0] for l in open(out) if l.startswith(">")]
        # the reference's rule: first record = first title line at or after byte B*idx; it keeps reading while the position
        # after the previous record is <= B*(idx+1)
        want, pos = [], B * idx
        first = next((k for k, st in enumerate(starts) if st >= pos), None)
        k = first
        while k is not None and k < len(starts):
            if (k > first and starts[k] > B * (idx + 1)) or (k == first and pos > B * (idx + 1)):
                break
            want.append(names[k])
            k += 1
        assert got == want, (idx, got, want)
        seen += got
    assert set(seen) == set(names)
