"""Deterministic synthetic data for tests and bench: SILVA-like reference MSAs,
full-length / V4 queries, and small random family cases. Pure numpy; no reference code involved.

An MSA is (masks u8[], cols u32[], off u64[N+1], W) with SINA's 4-bit IUPAC masks
(A=1 G=2 C=4 U=8, +16 lowercase; reference src/aligned_base.h:38-52).
"""
import numpy as np

AMBIG = np.array([3, 12, 10, 5, 6, 9, 14, 11, 13, 7, 15], np.uint8)  # R Y K M S W B D H V N
MASK2RNA = np.frombuffer(b".AGRCMSVUWKDYHBN.agrcmsvuwkdyhbn", np.uint8)


def _core_columns(rng, L, W):
    """L sorted columns with spacing >= 2 (so column+1 is always free for an insertion), heavy-tailed gaps
    so that the alignment has conserved blocks and long gap stretches like a real rRNA MSA."""
    x = rng.random(L) ** 8
    budget = W - 2 * L - 8
    assert budget > 0, "W too small for L"
    g = 2 + np.floor(x * (budget / x.sum())).astype(np.int64)
    cols = np.cumsum(g) - g[0] + 2
    assert cols[-1] + 2 < W
    return cols.astype(np.uint32)


def _mutate(rng, seqs, rate):
    """substitute each base with probability rate (per row or scalar) by a different base code"""
    n, L = seqs.shape
    r = np.broadcast_to(np.asarray(rate, np.float64).reshape(-1, 1), (n, 1))
    m = rng.random((n, L)) < r
    shift = rng.integers(1, 4, (n, L), dtype=np.uint8)
    return np.where(m, (seqs + shift) & 3, seqs).astype(np.uint8)


class Tree:
    """root -> phyla (25%) -> genera (8%) -> leaves (0-5%) over a fixed column map."""

    def __init__(self, n_refs, W=50000, L=1500, seed=20260117, leaves_per_genus=100, genera_per_phylum=10):
        self.rng = np.random.default_rng(seed)
        rng = self.rng
        self.W, self.L = W, L
        self.core = _core_columns(rng, L, W)
        self.n_genera = max(1, -(-n_refs // leaves_per_genus))
        n_phyla = max(1, -(-self.n_genera // genera_per_phylum))
        root = rng.integers(0, 4, (1, L), dtype=np.uint8)
        phyla = _mutate(rng, np.repeat(root, n_phyla, 0), 0.25)
        self.genus_phylum = np.arange(self.n_genera) % n_phyla
        self.genera = _mutate(rng, phyla[self.genus_phylum], 0.08)

    def leaves(self, genus_ids, rng, sub_lo=0.0, sub_hi=0.05, indel=0.01, ambig=0.001, trunc_frac=0.1, trunc_max=100,
               lo=0, hi=None):
        """Generate leaves of the given genera. Returns (masks, cols, off). [lo,hi) restricts to a core range."""
        hi = self.L if hi is None else hi
        n = len(genus_ids)
        L = self.L
        seqs = _mutate(rng, self.genera[genus_ids], rng.uniform(sub_lo, sub_hi, n))
        keep = rng.random((n, L)) >= indel / 2                     # deletions: skip the column
        ins = rng.random((n, L)) < indel / 2                       # insertions: use the free neighbour column
        ins_base = rng.integers(0, 4, (n, L), dtype=np.uint8)
        # truncation of some rows (partial sequences; still >= L - 2*trunc_max bases)
        start = np.where(rng.random(n) < trunc_frac, rng.integers(0, trunc_max + 1, n), 0)
        stop = L - np.where(rng.random(n) < trunc_frac, rng.integers(0, trunc_max + 1, n), 0)
        pos = np.arange(L)[None, :]
        inr = (pos >= np.maximum(start[:, None], lo)) & (pos < np.minimum(stop[:, None], hi))
        keep &= inr
        ins &= inr
        masks2 = np.empty((n, L, 2), np.uint8)
        masks2[:, :, 0] = 1 << seqs
        masks2[:, :, 1] = 1 << ins_base
        amb = rng.random((n, L)) < ambig
        masks2[:, :, 0] = np.where(amb, AMBIG[rng.integers(0, len(AMBIG), (n, L))], masks2[:, :, 0])
        valid = np.stack([keep, ins], 2)
        cols2 = np.stack([self.core, self.core + 1], 1)[None, :, :]
        cols2 = np.broadcast_to(cols2, (n, L, 2))
        cnt = valid.reshape(n, -1).sum(1)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum(cnt)
        return masks2[valid], cols2[valid].astype(np.uint32), off


def synth_msa(n_refs, W=50000, L=1500, seed=20260117, chunk=20000):
    """Reference MSA of n_refs rows. Returns (Tree, masks, cols, off)."""
    t = Tree(n_refs, W, L, seed)
    rng = np.random.default_rng(seed + 1)
    genus = np.arange(n_refs) % t.n_genera
    rng.shuffle(genus)
    ms, cs, offs = [], [], [np.zeros(1, np.uint64)]
    base = 0
    for a in range(0, n_refs, chunk):
        m, c, o = t.leaves(genus[a:a + chunk], rng)
        ms.append(m); cs.append(c); offs.append(o[1:] + np.uint64(base))
        base += int(o[-1])
    return t, np.concatenate(ms), np.concatenate(cs), np.concatenate(offs)


def synth_queries(tree, n_queries, kind="full", seed=7):
    """Held-out leaves with 1-3 % extra noise; 'v4' keeps the core range that maps to E. coli 515-806.
    Returns (qmasks, qoff) -- unaligned (columns dropped)."""
    rng = np.random.default_rng(seed)
    genus = rng.integers(0, tree.n_genera, n_queries)
    lo, hi = (0, None) if kind == "full" else (int(tree.L * 515 / 1542), int(tree.L * 806 / 1542))
    out_m, out_off = [], [np.zeros(1, np.uint64)]
    base = 0
    for a in range(0, n_queries, 20000):
        m, _, o = tree.leaves(genus[a:a + 20000], rng, sub_lo=0.01, sub_hi=0.08, indel=0.02, ambig=0.0005,
                              trunc_frac=0.0, lo=lo, hi=hi)
        out_m.append(m); out_off.append(o[1:] + np.uint64(base))
        base += int(o[-1])
    return np.concatenate(out_m), np.concatenate(out_off)


# ------------------------------------------------------------------ small random cases (parity tests)
def random_case(rng, F=None, L=None, sub=None, indel=None, iupac=0.01, lowercase=0.0, wfac=None, overhang_p=0.3):
    """One small family + query as strings: (rows[F] each of width W, query). Mirrors the survey's probe
    space: F 1-12 rows, L 20-220, W up to 4L, substitution/indel 2-32 %, IUPAC codes, partial rows,
    query overhangs."""
    F = F or int(rng.integers(1, 13))
    L = L or int(rng.integers(20, 221))
    sub = sub if sub is not None else float(rng.choice([0.02, 0.05, 0.1, 0.2, 0.32]))
    indel = indel if indel is not None else float(rng.choice([0.02, 0.05, 0.1, 0.2]))
    wfac = wfac or float(rng.uniform(1.0, 4.0))
    W = max(L + 2, int(L * wfac))
    core = np.sort(rng.choice(W, L, replace=False))
    root = rng.integers(0, 4, L)
    letters = "AGCU"
    amb = "RYKMSWBDHVN"

    def derive(allow_ins_cols):
        row = {}
        seq = []
        a = int(rng.integers(0, max(1, L // 4))) if rng.random() < 0.3 else 0
        b = L - (int(rng.integers(0, max(1, L // 4))) if rng.random() < 0.3 else 0)
        for i in range(a, b):
            if rng.random() < indel / 2:
                continue
            c = root[i] if rng.random() >= sub else int(rng.integers(0, 4))
            ch = letters[c]
            if rng.random() < iupac:
                ch = amb[int(rng.integers(0, len(amb)))]
            if rng.random() < lowercase:
                ch = ch.lower()
            row[int(core[i])] = ch
            seq.append(ch)
            if rng.random() < indel / 2:
                ch2 = letters[int(rng.integers(0, 4))]
                seq.append(ch2)
                col = int(core[i]) + 1
                if allow_ins_cols and col < W and (i + 1 >= L or col < core[i + 1]):
                    row[col] = ch2
                elif allow_ins_cols:
                    seq.pop()
        return row, "".join(seq)

    rows = []
    for _ in range(F):
        row, _s = derive(True)
        if not row:
            row = {int(core[0]): "A"}
        s = ["-"] * W
        for c, ch in row.items():
            s[c] = ch
        rows.append("".join(s))
    _r, q = derive(False)
    if rng.random() < overhang_p:
        q = "".join(letters[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 8)))) + q
    if rng.random() < overhang_p:
        q = q + "".join(letters[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 8))))
    if len(q) < 2:
        q = q + "AG"
    return rows, q
