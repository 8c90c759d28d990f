"""The CPU oracle against (1) the reference test-suite's own known answers and (2) the golden
vectors produced by the real reference (tests/golden/make_golden.py). Runs without /root/reference."""
import pytest

import golden_util
from oracle import oracle
from oracle.oracle import BACK

# R=A|G ... the reference's wildcard fixtures, /root/reference/tests/test_align.py:36-52
WILDCARD_SEQUENCES = [
    'CCCATTGATC', 'CCCRTTRATC', 'YCCATYGATC', 'CSSATTSATC', 'CCCWWWGATC', 'CCCATKKATC', 'CCMATTGMTC',
    'BCCATTBABC', 'BCCATTBABC', 'CCCDTTDADC', 'CHCATHGATC', 'CVCVTTVATC', 'CCNATNGATC', 'CCCNTTNATC',
]


def test_kat_polya():                         # tests/test_align.py:24-29
    s = 'AAAAAAAAAAAAAAAAA'
    t = 'ACAGAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA'
    assert oracle.locate(s, t, 0.0, BACK) == (0, len(s), 4, 4 + len(s), len(s), 0)


def test_kat_smoke_cases():                   # tests/test_align.py:13-21
    oracle.locate('CTCCAGCTTAGACATATC', 'CC', 0.1, BACK)
    oracle.locate('GCTTAGACATATC', 'CAA', 1.0, BACK)


def test_kat_compare_prefixes():              # tests/test_align.py:55-84
    cp = oracle.compare_prefixes
    assert cp('AAXAA', 'AAAAATTTTTTTTT') == (0, 5, 0, 5, 4, 1)
    assert cp('AANAA', 'AACAATTTTTTTTT', wildcard_ref=True) == (0, 5, 0, 5, 5, 0)
    assert cp('XAAAAA', 'AAAAATTTTTTTTT') == (0, 6, 0, 6, 4, 2)
    a = WILDCARD_SEQUENCES[0]
    for s in WILDCARD_SEQUENCES:
        r = s + 'GCCAGGGTTGATTCGGCTGATCTGGCCG'
        assert cp(a, r, wildcard_query=True) == (0, 10, 0, 10, 10, 0)
        assert cp(r, a, wildcard_ref=True) == (0, 10, 0, 10, 10, 0)
    for s in WILDCARD_SEQUENCES:
        r = s + 'GCCAGGG'
        assert cp(s, r) == (0, 10, 0, 10, 10, 0)
        assert cp(r, s, wildcard_ref=True, wildcard_query=True) == (0, 10, 0, 10, 10, 0)
    r = WILDCARD_SEQUENCES[0] + 'GCCAGG'
    for wr in (False, True):
        for wq in (False, True):
            assert cp('CCCXTTXATC', r, wildcard_ref=wr, wildcard_query=wq) == (0, 10, 0, 10, 8, 2)


def test_kat_compare_suffixes():              # tests/test_align.py:87-95
    cs = oracle.compare_suffixes
    assert cs('AAXAA', 'TTTTTTTAAAAA') == (0, 5, 7, 12, 4, 1)
    assert cs('AANAA', 'TTTTTTTAACAA', wildcard_ref=True) == (0, 5, 7, 12, 5, 0)
    assert cs('AAAAAX', 'TTTTTTTAAAAA') == (0, 6, 6, 12, 4, 2)


def test_kat_wildcards():                     # tests/test_align.py:98-127
    r = 'CATCTGTCC' + WILDCARD_SEQUENCES[0] + 'GCCAGGGTTGATTCGGCTGATCTGGCCG'
    for a in WILDCARD_SEQUENCES:
        assert oracle.locate(a, r, 0.0, BACK, wildcard_ref=True) == (0, 10, 9, 19, 10, 0)
    assert oracle.locate('CCCXTTXATC', r, 0.0, BACK, wildcard_ref=True) is None
    a = WILDCARD_SEQUENCES[0]
    for s in WILDCARD_SEQUENCES:
        rr = 'CATCTGTCC' + s + 'GCCAGGGTTGATTCGGCTGATCTGGCCG'
        assert oracle.locate(a, rr, 0.0, BACK, wildcard_query=True) == (0, 10, 9, 19, 10, 0)
    for a in WILDCARD_SEQUENCES:
        for s in WILDCARD_SEQUENCES:
            rr = 'CATCTGTCC' + s + 'GCCAGGGTTGATTCGGCTGATCTGGCCG'
            assert oracle.locate(a, rr, 0.0, BACK, wildcard_ref=True, wildcard_query=True) == (0, 10, 9, 19, 10, 0)


def test_kat_no_match():                      # tests/test_align.py:130-132
    assert oracle.locate('CTGATCTGGCCG', 'AAAAGGG', 0.1, BACK) is None


def test_kat_match_probability():             # tests/test_align.py:135-153
    import math
    f = oracle.RandomMatchProbability()
    assert f.factorial(0) == 1 and f.factorial(1) == 1 and f.factorial(3) == 6
    assert f.factorial(27) == math.factorial(27) and f.factorial(150) == math.factorial(150)
    i3 = (120 / (6 * 2)) * (0.25 ** 3) * (0.75 ** 2)
    i4 = (120 / 24) * (0.25 ** 4) * 0.75
    i5 = 0.25 ** 5
    assert abs(f(3, 5) - (i3 + i4 + i5)) < 0.0001


def test_kat_insert_align():                  # tests/test_align.py:156-179
    a1, a2 = 'TTAGACATATGG', 'CAGTGGAGTATA'
    al = oracle.OracleInsertAligner(a1, a2)
    _, m1, m2 = al.match_insert('AGTCGAGCCCATTGCAGACT' + a1[0:10], 'AGTCTGCAATGGGCTCGACT' + a2[0:10])
    assert m1[2] == 20 and m1[1] - m1[0] == 10 and m2[2] == 20 and m2[1] - m2[0] == 10
    al = oracle.OracleInsertAligner('TTAGACATAT', 'CAGTGGAGTA')
    _, m1, m2 = al.match_insert('GACAGGCCGTTTGAATGTTGACGGGATGTT', 'CATCCCGTCAACATTCAAACGGCCTGTCCA')
    assert m1[2] == 28 and m1[1] - m1[0] == 2 and m2[2] == 28 and m2[1] - m2[0] == 2


def test_kat_multi_aligner():                 # tests/test_align.py:195-234
    ms = oracle.multi_locate('AGAGATCAGATGACAGATC', 'GATCA', 0, min_overlap=3)
    assert len(ms) == 2
    ms.sort(key=lambda x: x[4], reverse=True)
    assert ms[0] == (3, 8, 0, 5, 5, 0) and ms[1] == (15, 19, 0, 4, 4, 0)
    ms = oracle.multi_locate('GATATCAGATGACAGATCAGAGATCAGAT', 'GAGATCAGATGA', 0.1, min_overlap=10)
    assert len(ms) == 2
    ms.sort(key=lambda x: x[5])
    assert ms[0] == (19, 29, 0, 10, 10, 0) and ms[1] == (0, 12, 0, 12, 11, 1)


def test_kat_issue_80():                      # tests/test_adapters.py:44-68 (indel tie-break)
    ad = oracle.OracleAdapter("TCGTATGCCGTCTTC", BACK, max_error_rate=0.2, min_overlap=3,
                              read_wildcards=False, adapter_wildcards=False)
    res = ad.match_to("TCGTATGCCCTCC")
    assert res[5] == 3 and res[0] == 0 and res[1] == 15


def test_kat_linked():                        # tests/test_adapters.py:119-125
    fa, ba = oracle.OracleAdapter('AAAA', oracle.PREFIX), oracle.OracleAdapter('TTTT', BACK)
    fm, bm = oracle.linked_match_to(fa, ba, 'AAAACCCCCTTTT')
    assert 'AAAACCCCCTTTT'[fm[3]:][:bm[2]] == 'CCCCC'


# ---- golden vectors from the real reference ----------------------------------------------

def test_golden_locate():
    cases = golden_util.load("locate")
    found = 0
    for c in cases:
        got = oracle.locate(c["reference"], c["query"], c["max_error_rate"], c["flags"], c["wildcard_ref"],
                            c["wildcard_query"], c["min_overlap"], c["indel_cost"])
        exp = None if c["expect"] is None else tuple(c["expect"])
        assert got == exp, c
        found += exp is not None
    assert len(cases) == 5000 and found > 1500


def test_golden_multi_locate():
    for c in golden_util.load("multi_locate"):
        got = oracle.multi_locate(c["reference"], c["query"], c["max_error_rate"], c["flags"], c["min_overlap"])
        exp = None if c["expect"] is None else [tuple(t) for t in c["expect"]]
        assert got == exp, c


def test_golden_match_to():
    rmp = oracle.RandomMatchProbability()
    n = 0
    for c in golden_util.load("match_to"):
        ad = oracle.OracleAdapter(c["sequence"], c["where"], match_probability=rmp, max_rmp=c["max_rmp"], **c["kw"])
        for read, exp in zip(c["reads"], c["expect"]):
            got = ad.match_to(read)
            assert (None if got is None else list(got)) == exp, (c, read)
            n += exp is not None
    assert n > 400


def test_golden_match_insert():
    n = 0
    for g in golden_util.load("match_insert"):
        al = oracle.OracleInsertAligner(g["adapter1"], g["adapter2"], **g["kw"])
        for (r1, r2), exp in zip(g["pairs"], g["expect"]):
            got = al.match_insert(r1, r2)
            if exp is None:
                assert got is None
                continue
            n += 1
            assert list(got[0]) == exp[0]
            for gm, em in ((got[1], exp[1]), (got[2], exp[2])):
                assert (None if gm is None else list(gm)) == (None if em is None else em[:6])
    assert n > 500


def test_golden_rmp():
    f = oracle.RandomMatchProbability()
    for matches, size, hx in golden_util.load("rmp"):
        assert f(matches, size) == float.fromhex(hx)
