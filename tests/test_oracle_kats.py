"""Known-answer vectors held by the reference's own unit tests, replayed on the C restatement
(and on the compiled reference when present)."""
import numpy as np
import pytest

from oracle import oracle as O

# src/unit_tests/kmer_test.cpp:51-54
TEST_SEQUENCE = "AGCTN" "AGCTAGCTN" "AGCTAGCTAGCTN"
TRY_K = [1, 2, 3, 4, 8]
CODE = {"A": 0, "G": 1, "C": 2, "T": 3}  # src/aligned_base.h:38-45

# src/unit_tests/kmer_test.cpp:61-87 (valid_k) and :89-115 (first_k), one string per k, '1' = true
VALID_K = ["111101111111101111111111110", "011100111111100111111111110", "001100011111100011111111110",
           "000100001111100001111111110", "000000000000100000000111110"]
FIRST_K = ["111100000000000000000000000", "011100000100000000000000000", "001100000110000000000000000",
           "000100000111000000000000000", "000000000000100000000011100"]
# src/unit_tests/kmer_test.cpp:273-304: (prefix_len, prefix) -> count; only the A-prefix (p_len 1) rows
# apply to the hot path's prefix_kmers(bases, k, 1, BASE_A): k=1 {A}: 6, k=2 {A}: 6


def kmer_value(i, k):
    """kmers_k table (src/unit_tests/kmer_test.cpp:122-163): value of the k-mer ending at position i"""
    v = 0
    for ch in TEST_SEQUENCE[i - k + 1:i + 1]:
        v = (v << 2) | CODE[ch]
    return v


@pytest.mark.parametrize("n", range(len(TRY_K)))
def test_kmer_generator_tables(orc, n):
    k = TRY_K[n]
    masks = O.encode(TEST_SEQUENCE)
    exp_all = [kmer_value(i, k) for i in range(len(TEST_SEQUENCE)) if VALID_K[n][i] == "1"]
    exp_uniq = [kmer_value(i, k) for i in range(len(TEST_SEQUENCE)) if FIRST_K[n][i] == "1"]
    # the sequence ends in N, so the dropped-last-k-mer rule (src/kmer.h:179-201) is not exercised here
    assert orc.kmers(masks, k, 0).tolist() == exp_all
    assert orc.kmers(masks, k, 1).tolist() == exp_uniq
    exp_pref = [v for v in exp_all if (v >> (2 * (k - 1))) == 0]
    assert orc.kmers(masks, k, 2).tolist() == exp_pref
    if k in (1, 2):
        assert len(exp_pref) == 6  # kmer_prefix_counts for prefix A
    seen, exp_upref = set(), []
    for i in range(len(TEST_SEQUENCE)):
        if VALID_K[n][i] == "1":
            v = kmer_value(i, k)
            if (v >> (2 * (k - 1))) == 0 and v not in seen:
                exp_upref.append(v)
            seen.add(v)
    assert orc.kmers(masks, k, 3).tolist() == exp_upref
    assert len(exp_upref) <= 1 or k > 2


def test_kmer_last_base_dropped(orc):
    """SURVEY Appendix B KAT: 8-mer string, k=4 -> 4 k-mers (not 5); prefix-A -> {27}"""
    m = O.encode("AGCTAGCA")
    assert orc.kmers(m, 4, 0).tolist() == [27, 108, 177, 198]
    assert orc.kmers(m, 4, 2).tolist() == [27]
    assert orc.kmers(O.encode("AGCT"), 4, 0).tolist() == []
    assert orc.kmers(np.zeros(0, np.uint8), 4, 0).tolist() == []


def test_kmer_tables_on_reference(ref):
    for n, k in enumerate(TRY_K):
        exp_all = [kmer_value(i, k) for i in range(len(TEST_SEQUENCE)) if VALID_K[n][i] == "1"]
        assert ref.kmers(TEST_SEQUENCE, k, 0).tolist() == exp_all


def test_base_encoding(orc):
    """src/unit_tests/aligned_base_test.cpp:43-196"""
    A, G, C_, U = (orc.char_to_mask(c) for c in "AGCU")
    assert (A, G, C_, U) == (1, 2, 4, 8)
    assert orc.char_to_mask("T") == U and orc.char_to_mask("!") == -1
    assert orc.mask_to_char(A) == "A" and orc.mask_to_char(U) == "U" and orc.mask_to_char(U, dna=True) == "T"
    assert orc.mask_to_char(A | 16) == "a" and orc.mask_to_char(C_ | 16) == "c"       # setLower_test
    assert orc.mask_to_char(orc.char_to_mask("t") & 15) == "U"                        # setUpper_test
    assert orc.char_to_mask("t") & 16 and not orc.char_to_mask("G") & 16              # isLower_test
    comp = lambda a, b: (orc.char_to_mask(a) & orc.char_to_mask(b) & 15) != 0          # comp_test
    assert comp("T", "U") and comp("T", "u") and comp("u", "T") and comp("t", "T") and comp("G", "G")
    assert not comp("T", "C") and not comp("G", "C") and not comp("t", "C")
    pop = lambda c: bin(orc.char_to_mask(c) & 15).count("1")                           # ambig_order_test
    assert (pop("T"), pop("M"), pop("D")) == (1, 2, 3)
    for c in "abcdghkmnrstuvwy":                                                       # cast_to_char_test
        m = orc.char_to_mask(c)
        assert m > 0 and orc.mask_to_char(m) == ("u" if c == "t" else c)


def test_cseq_strings(orc):
    """src/unit_tests/cseq_test.cpp:49-52,99-248: aligned string <-> (bases, columns) round trips"""
    rna = "AGCURYKMSWBDHVN"
    rna_aligned = "--A-G---CUR-YKM-S---WBD-HVN---"
    for s in (rna, rna_aligned, rna_aligned.lower(), rna + rna_aligned):
        msa = O.MSA.from_rows([s])
        m, c = msa.row(0)
        assert O.decode(m) == s.replace("-", "")
        assert O.render(m, c, msa.W) == s
    msa = O.MSA.from_rows([rna_aligned.lower()])                                       # test_dna
    m, c = msa.row(0)
    dna = "".join(orc.mask_to_char(int(x), dna=True) for x in m)
    assert dna == rna.lower().replace("u", "t")


def test_cseq_strings_on_reference(ref):
    rna_aligned = "--A-G---CUR-YKM-S---WBD-HVN---"
    assert ref.cseq_roundtrip(rna_aligned) == rna_aligned
    assert ref.cseq_roundtrip(rna_aligned.lower(), dna=True) == rna_aligned.lower().replace("u", "t")
    for c in "AGCUTRYKMSWBDHVNagcutrykmswbdhvn-.":
        assert ref.char_to_mask(c) == O._CHAR2MASK[ord(c)]
