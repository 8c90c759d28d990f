"""Parity of the CUDA path (through the C ABI / the reference-shaped Python mirror) with the CPU oracle
and the committed golden vectors from the real reference. Bar: BIT-EXACT on every field
(astart, astop, rstart, rstop, matches, errors, None-ness, winning adapter). Needs a GPU."""
import numpy as np
import pytest
from atropos_b200 import _abi as _abi_mod

import fuzzgen
import golden_util
from oracle import oracle

pytestmark = pytest.mark.gpu

T1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"
T2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT"
FIELDS = ("astart", "astop", "rstart", "rstop", "matches", "errors")


def _tup(rec):
    return tuple(int(rec[k]) for k in FIELDS)


def _check_locate_array(got, exp7):
    """got: MATCH_DTYPE array; exp7: int32 [n,7] from oracle.locate_batch."""
    from atropos_b200 import _abi
    found = exp7[:, 0] == 1
    assert np.array_equal(got["status"] == _abi.ATR_ST_MATCH, found)
    assert np.array_equal(got["status"] == _abi.ATR_ST_NONE, ~found)
    for c, k in enumerate(FIELDS):
        assert np.array_equal(got[k][found].astype(np.int64), exp7[found, c + 1].astype(np.int64)), k
    return int(found.sum())


def test_reference_known_answers():
    """the reference test-suite's KATs through the drop-in names (tests/test_align.py, tests/test_adapters.py)"""
    from atropos_b200.adapters import Adapter, BACK, LinkedAdapter
    from atropos_b200.align import Aligner, InsertAligner, MultiAligner, compare_prefixes, compare_suffixes, locate
    Aligner('CTCCAGCTTAGACATATC', 0.1, flags=BACK).locate('CC')
    Aligner('GCTTAGACATATC', 1.0, flags=BACK).locate('CAA')
    s, t = 'AAAAAAAAAAAAAAAAA', 'ACAGAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA'
    assert locate(s, t, 0.0, BACK) == (0, len(s), 4, 4 + len(s), len(s), 0)
    assert compare_prefixes('AAXAA', 'AAAAATTTTTTTTT') == (0, 5, 0, 5, 4, 1)
    assert compare_prefixes('AANAA', 'AACAATTTTTTTTT', wildcard_ref=True) == (0, 5, 0, 5, 5, 0)
    assert compare_prefixes('XAAAAA', 'AAAAATTTTTTTTT') == (0, 6, 0, 6, 4, 2)
    assert compare_suffixes('AAXAA', 'TTTTTTTAAAAA') == (0, 5, 7, 12, 4, 1)
    assert compare_suffixes('AANAA', 'TTTTTTTAACAA', wildcard_ref=True) == (0, 5, 7, 12, 5, 0)
    assert compare_suffixes('AAAAAX', 'TTTTTTTAAAAA') == (0, 6, 6, 12, 4, 2)
    W = ['CCCATTGATC', 'CCCRTTRATC', 'YCCATYGATC', 'CSSATTSATC', 'CCCWWWGATC', 'CCCATKKATC', 'CCMATTGMTC', 'BCCATTBABC',
         'CCCDTTDADC', 'CHCATHGATC', 'CVCVTTVATC', 'CCNATNGATC', 'CCCNTTNATC']
    r = 'CATCTGTCC' + W[0] + 'GCCAGGGTTGATTCGGCTGATCTGGCCG'
    for a in W:
        assert locate(a, r, 0.0, BACK, wildcard_ref=True) == (0, 10, 9, 19, 10, 0)
    assert locate('CCCXTTXATC', r, 0.0, BACK, wildcard_ref=True) is None
    for sq in W:
        rr = 'CATCTGTCC' + sq + 'GCCAGGGTTGATTCGGCTGATCTGGCCG'
        assert locate(W[0], rr, 0.0, BACK, wildcard_query=True) == (0, 10, 9, 19, 10, 0)
        assert locate(W[3], rr, 0.0, BACK, wildcard_ref=True, wildcard_query=True) == (0, 10, 9, 19, 10, 0)
    assert locate('CTGATCTGGCCG', 'AAAAGGG', 0.1, BACK) is None
    # issue 80: the indel tie-break (tests/test_adapters.py:44-68)
    ad = Adapter("TCGTATGCCGTCTTC", BACK, max_error_rate=0.2, min_overlap=3, read_wildcards=False, adapter_wildcards=False)
    res = ad.match_to("TCGTATGCCCTCC")
    assert res.errors == 3 and res.astart == 0 and res.astop == 15
    # insert aligner (tests/test_align.py:156-179)
    a1, a2 = 'TTAGACATATGG', 'CAGTGGAGTATA'
    _, m1, m2 = InsertAligner(a1, a2).match_insert('AGTCGAGCCCATTGCAGACT' + a1[0:10], 'AGTCTGCAATGGGCTCGACT' + a2[0:10])
    assert (m1.rstart, m1.length, m2.rstart, m2.length) == (20, 10, 20, 10)
    _, m1, m2 = InsertAligner('TTAGACATAT', 'CAGTGGAGTA').match_insert('GACAGGCCGTTTGAATGTTGACGGGATGTT',
                                                                       'CATCCCGTCAACATTCAAACGGCCTGTCCA')
    assert (m1.rstart, m1.length, m2.rstart, m2.length) == (28, 2, 28, 2)
    # MultiAligner (tests/test_align.py:195-234)
    ms = MultiAligner(max_error_rate=0, min_overlap=3).locate('AGAGATCAGATGACAGATC', 'GATCA')
    assert sorted(ms, key=lambda x: -x[4]) == [(3, 8, 0, 5, 5, 0), (15, 19, 0, 4, 4, 0)]
    ms = MultiAligner(max_error_rate=0.1, min_overlap=10).locate('GATATCAGATGACAGATCAGAGATCAGAT', 'GAGATCAGATGA')
    assert sorted(ms, key=lambda x: x[5]) == [(19, 29, 0, 10, 10, 0), (0, 12, 0, 12, 11, 1)]
    # linked adapter (tests/test_adapters.py:119-125)
    lm = LinkedAdapter('AAAA', 'TTTT').match_to('AAAACCCCCTTTT')
    assert 'AAAACCCCCTTTT'[lm.front_match.rstop:][:lm.back_match.rstart] == 'CCCCC'
    with pytest.raises(ValueError):
        Aligner('ACGT', 0.1).min_overlap = 0


def test_golden_locate():
    """5000 reference-generated cases; grouped so that each adapter configuration is one GPU batch"""
    from atropos_b200.align import Aligner
    cases = golden_util.load("locate")
    groups = {}
    for c in cases:
        key = (c["reference"], c["max_error_rate"], c["flags"], c["wildcard_ref"], c["wildcard_query"], c["min_overlap"],
               c["indel_cost"])
        groups.setdefault(key, []).append(c)
    from atropos_b200 import _abi
    for key, cs in groups.items():
        al = Aligner(key[0], key[1], key[2], key[3], key[4], key[5], key[6])
        res = al.locate_batch([c["query"] for c in cs])
        for c, rec in zip(cs, res):
            got = None if rec["status"] == _abi.ATR_ST_NONE else _tup(rec)
            assert got == (None if c["expect"] is None else tuple(c["expect"])), c


def test_golden_match_to_insert_multi():
    from atropos_b200 import _abi
    from atropos_b200.adapters import Adapter
    from atropos_b200.align import InsertAligner, MultiAligner
    from atropos_b200.util import RandomMatchProbability
    rmp = RandomMatchProbability()
    for c in golden_util.load("match_to"):
        ad = Adapter(c["sequence"], c["where"], match_probability=rmp, max_rmp=c["max_rmp"], **c["kw"])
        res = ad.match_to_batch(c["reads"])
        for read, exp, rec in zip(c["reads"], c["expect"], res):
            m = ad.match_from_record(rec)
            assert (None if m is None else list(m.fields()) + [bool(m.front)]) == exp, (c, read)
    for g in golden_util.load("match_insert"):
        ia = InsertAligner(g["adapter1"], g["adapter2"], **g["kw"])
        res = ia.match_insert_batch([p[0] for p in g["pairs"]], [p[1] for p in g["pairs"]])
        for exp, rec in zip(g["expect"], res):
            got = InsertAligner.result_from_record(rec)
            if exp is None:
                assert got is None
            else:
                assert list(got[0]) == exp[0]
                for gm, em in ((got[1], exp[1]), (got[2], exp[2])):
                    assert (None if gm is None else list(gm.fields())) == (None if em is None else em[:6])
    for c in golden_util.load("multi_locate")[:200]:
        got = MultiAligner(c["max_error_rate"], c["flags"], c["min_overlap"]).locate(c["reference"], c["query"])
        assert got == (None if c["expect"] is None else [tuple(t) for t in c["expect"]])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_locate_fuzz_batches(seed):
    """random adapter configurations x ragged batches, all flag sets / indel costs / modes, vs the C oracle"""
    from atropos_b200.align import Aligner
    rng = np.random.default_rng(900 + seed)
    total = 0
    for _ in range(40):
        wild = rng.random() < 0.3
        m = int(rng.integers(1, 65)) if rng.random() < 0.85 else int(rng.integers(65, 200))
        ad = fuzzgen.rand_seq(rng, m, "ACGTACGTACGTNRYSWKMBDHVX" if wild else "ACGT")
        wr, wq = [(False, False), (True, False), (False, True), (True, True)][int(rng.integers(1, 4)) if wild else 0]
        rate = float(rng.choice([0.0, 0.05, 0.1, 0.15, 0.2, 0.3, 0.5, 1.0]))
        flags = int(rng.choice([14, 14, 11, 8, 2, 15, 9, 0, 5, 7]))
        mo, ic = int(rng.choice([1, 3, 5, 10])), int(rng.choice([1, 1, 2, 3, 100000]))
        proj = "".join(ch if ch in "ACGT" else "ACGT"[int(rng.integers(0, 4))] for ch in ad)
        reads = []
        for _ in range(600):
            r = fuzzgen.read_with_adapter(rng, proj, int(rng.integers(0, 230)))
            if rng.random() < 0.05:
                arr = list(r)
                for i in range(len(arr)):
                    if rng.random() < 0.1:
                        arr[i] = "acgtnNRYU."[int(rng.integers(0, 10))]
                r = "".join(arr)
            reads.append(r)
        blob = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
        offs = np.zeros(len(reads) + 1, dtype=np.int64)
        np.cumsum([len(r) for r in reads], out=offs[1:])
        exp = oracle.locate_batch(ad, blob, offs, rate, flags, wr, wq, mo, ic)
        got = Aligner(ad, rate, flags, wr, wq, mo, ic).locate_batch((blob, offs))
        total += _check_locate_array(got, exp)
    assert total > 2000


def test_headline_config_se150():
    """BASELINE config 2 at a size the oracle finishes in seconds: 300 k synthetic 150 nt reads, TruSeq, 0.1"""
    from atropos_b200 import synth
    from atropos_b200.adapters import Adapter, BACK
    n, L = 300000, 150
    reads = synth.synth_se(n, L, seed=synth.seed_for(2), device="cpu").numpy().reshape(-1)
    offs = np.arange(n + 1, dtype=np.int64) * L
    for adapter, mo in ((T1, 3), ("AGATCGGAAGAGC", 3)):
        got = Adapter(adapter, BACK, 0.1, mo).match_to_batch((reads, offs))
        exp = oracle.locate_batch(adapter, reads, offs, 0.1, oracle.BACK, False, False, mo, 1)
        # match_to's post-filter never rejects what locate accepts for BACK adapters without max_rmp
        hits = _check_locate_array(got, exp)
        assert hits > n // 4


def test_match_to_semantics_batches():
    from atropos_b200 import _abi
    from atropos_b200.adapters import Adapter, ANYWHERE, BACK, FRONT, PREFIX, SUFFIX
    from atropos_b200.util import RandomMatchProbability
    rng = np.random.default_rng(77)
    rmp, rmp_o = RandomMatchProbability(), oracle.RandomMatchProbability()
    found = 0
    for _ in range(120):
        where = [BACK, FRONT, ANYWHERE, PREFIX, SUFFIX][int(rng.integers(0, 5))]
        wild = rng.random() < 0.3
        seq = fuzzgen.rand_seq(rng, int(rng.integers(3, 70)), "ACGTACGTACGTNRY" if wild else "ACGT")
        kw = dict(max_error_rate=float(rng.choice([0.0, 0.1, 0.12, 0.2])), min_overlap=int(rng.choice([1, 3, 5])),
                  read_wildcards=bool(rng.random() < 0.3), adapter_wildcards=bool(rng.random() < 0.7),
                  indels=bool(rng.random() < 0.7), indel_cost=int(rng.choice([1, 1, 3])))
        max_rmp = [None, 1e-6, 1e-3][int(rng.integers(0, 3))]
        mine = Adapter(seq, where, match_probability=rmp, max_rmp=max_rmp, **kw)
        orc = oracle.OracleAdapter(seq, where, match_probability=rmp_o, max_rmp=max_rmp, **kw)
        proj = "".join(ch if ch in "ACGT" else "ACGT"[int(rng.integers(0, 4))] for ch in seq)
        reads = []
        for _ in range(100):
            r = fuzzgen.read_with_adapter(rng, proj if rng.random() < 0.7 else seq, int(rng.integers(0, 160)), n_rate=0.02)
            reads.append(r.lower() if rng.random() < 0.15 else r)
        res = mine.match_to_batch(reads)
        for r, rec in zip(reads, res):
            exp = orc.match_to(r)
            m = mine.match_from_record(rec)
            assert (None if m is None else m.fields() + (m.front,)) == exp, (seq, where, kw, max_rmp, r)
            found += exp is not None
    assert found > 1500


def test_panel_times_linked():
    """config 4's shapes: best-of-N panel (_best_match), `times` rounds via windows, linked adapters"""
    from atropos_b200 import _abi
    from atropos_b200.adapters import Adapter, BACK, FRONT, LinkedAdapter, PREFIX
    from atropos_b200.modifiers import AdapterCutter
    rng = np.random.default_rng(88)
    specs = [(T1, BACK), (T2, BACK), ("TGGAATTCTCGGGTGCCAAGG", BACK), ("GTTCAGAGTTCTACAGTCCGACGATC", PREFIX),
             ("ACACTCTTTCCCTACACGACGCTCTTCCGATCT", PREFIX), ("AATGATACGGCGACCACCGA", FRONT)]
    mine = [Adapter(s, w) for s, w in specs]
    orc = [oracle.OracleAdapter(s, w) for s, w in specs]
    reads = []
    for _ in range(4000):
        s, w = specs[int(rng.integers(0, len(specs)))]
        body = fuzzgen.read_with_adapter(rng, s, 150)
        if w != BACK and rng.random() < 0.7:
            body = (fuzzgen.mutate(rng, s, 0.03, 0.01, 0.01) + body)[:150]
        reads.append(body)
    cutter = AdapterCutter(mine, times=2)
    rounds = cutter.match_rounds_batch(reads)
    nhit = n2 = 0
    for i, r in enumerate(reads):
        cur = r
        for t in range(2):
            exp = oracle.best_match(orc, cur) if len(cur) else None
            rec = rounds[t][i] if t < len(rounds) else None
            if exp is None:
                assert rec is None or rec["status"] == _abi.ATR_ST_NONE, (i, t, r)
                break
            assert rec["status"] == _abi.ATR_ST_MATCH and int(rec["adapter"]) == exp[0] and _tup(rec) == exp[1][:6], (i, t, r, exp, rec)
            nhit += t == 0
            n2 += t == 1
            cur = cur[exp[1][3]:] if exp[1][6] else cur[:exp[1][2]]
    assert nhit > 2000 and n2 > 20
    front, back = "GTTCAGAGTTCTACAGTCCGACGATC", "TGGAATTCTCGGGTGCCAAGG"
    la = LinkedAdapter(front, back)
    fa, ba = oracle.OracleAdapter(front, oracle.PREFIX), oracle.OracleAdapter(back, oracle.BACK)
    lreads = []
    for _ in range(2000):
        f = fuzzgen.mutate(rng, front, 0.03, 0.01, 0.01) if rng.random() < 0.7 else fuzzgen.rand_seq(rng, 26)
        lreads.append((f + fuzzgen.rand_seq(rng, int(rng.integers(0, 60))) + fuzzgen.mutate(rng, back, 0.03, 0.01, 0.01) +
                       fuzzgen.rand_seq(rng, 30))[:100])
    fr, bk = la.match_to_batch(lreads)
    for i, r in enumerate(lreads):
        exp = oracle.linked_match_to(fa, ba, r)
        if exp is None:
            assert fr[i]["status"] == _abi.ATR_ST_NONE and bk[i]["status"] == _abi.ATR_ST_NONE
        else:
            assert _tup(fr[i]) == exp[0][:6]
            assert (None if bk[i]["status"] == _abi.ATR_ST_NONE else _tup(bk[i])) == (None if exp[1] is None else exp[1][:6])


def test_insert_config_pe150():
    """BASELINE config 3 shape: 2x150 synthetic pairs, TruSeq R1/R2, rate 0.1, vs the Python oracle"""
    from atropos_b200 import synth
    from atropos_b200.align import InsertAligner
    n, L = 20000, 150
    r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(3), device="cpu")
    r1, r2 = r1.numpy(), r2.numpy()
    offs = np.arange(n + 1, dtype=np.int64) * L
    kw = dict(max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1)
    res = InsertAligner(T1, T2, **kw).match_insert_batch((r1.reshape(-1), offs), (r2.reshape(-1), offs))
    orc = oracle.OracleInsertAligner(T1, T2, **kw)
    nm = 0
    for i in range(n):
        exp = orc.match_insert(bytes(r1[i]).decode(), bytes(r2[i]).decode())
        got = InsertAligner.result_from_record(res[i])
        if exp is None:
            assert got is None, i
        else:
            nm += 1
            assert got[0] == exp[0] and (got[1].fields() if got[1] else None) == exp[1] and \
                (got[2].fields() if got[2] else None) == exp[2], (i, got, exp)
    assert nm > n // 4


def test_insert_config_pe300():
    """BASELINE config 5 shape: 1 M synthetic 2x300 pairs, error rate 0.15 (k = 45), every pair vs the Python oracle; also the
    per-read fallback adapters of insert mode (max_rmp 1e-6, min_overlap 1, indel cost 3) at 300 nt"""
    from atropos_b200 import synth
    from atropos_b200.adapters import Adapter, BACK
    from atropos_b200.align import InsertAligner
    from atropos_b200.util import RandomMatchProbability
    n, L = 1_000_000, 300
    r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(5), device="cuda", sub=0.02)
    r1, r2 = r1.cpu().numpy(), r2.cpu().numpy()
    offs = np.arange(n + 1, dtype=np.int64) * L
    kw = dict(max_insert_mismatch_frac=0.15, max_adapter_mismatch_frac=0.15)
    res = InsertAligner(T1, T2, **kw).match_insert_batch((r1.reshape(-1), offs), (r2.reshape(-1), offs))
    nm = _pe_check_all(res, r1, r2, kw)                       # all 1 M pairs, forked oracle workers
    assert nm > n // 4
    rmp, rmp_o = RandomMatchProbability(), oracle.RandomMatchProbability()
    akw = dict(max_error_rate=0.15, min_overlap=1, indel_cost=3, max_rmp=1e-6)
    mine = Adapter(T2, BACK, match_probability=rmp, **akw)
    o_ad = oracle.OracleAdapter(T2, oracle.BACK, match_probability=rmp_o, **akw)
    rec = mine.match_to_batch((r2.reshape(-1), offs))
    for i in range(0, 30000, 5):
        exp = o_ad.match_to(bytes(r2[i]).decode())
        got = None if rec[i]["status"] == _abi_mod.ATR_ST_NONE else tuple(int(rec[i][k]) for k in
                                                                         ("astart", "astop", "rstart", "rstop", "matches", "errors"))
        assert got == (None if exp is None else tuple(exp[:6])), i


def test_insert_fuzz_modes_and_escapes():
    from atropos_b200 import _abi
    from atropos_b200.align import InsertAligner
    rng = np.random.default_rng(99)
    cfgs = [dict(), dict(max_insert_mismatch_frac=0.15, max_adapter_mismatch_frac=0.15),
            dict(max_insert_mismatch_frac=0.3, max_adapter_mismatch_frac=0.2, min_insert_overlap=10, adapter_wildcards=False),
            dict(read_wildcards=True), dict(read_wildcards=True, adapter_wildcards=False)]
    for ci, kw in enumerate(cfgs):
        pairs = []
        for r1, r2 in fuzzgen.insert_pairs(600 + ci, 1500, T1, T2, lengths=(50, 100, 150, 300), err=0.03):
            if rng.random() < 0.05:
                pos = int(rng.integers(0, len(r2)))
                r2 = r2[:pos] + "aX.n"[int(rng.integers(0, 4))] + r2[pos + 1:]
            pairs.append((r1, r2))
        ia = InsertAligner(T1, T2, **kw)
        res = ia.match_insert_batch([p[0] for p in pairs], [p[1] for p in pairs])
        orc = oracle.OracleInsertAligner(T1, T2, **kw)
        for (r1, r2), rec in zip(pairs, res):
            try:
                exp = orc.match_insert(r1, r2)
            except KeyError:
                assert int(rec["insert"]["status"]) == _abi.ATR_ST_KEYERROR
                continue
            got = InsertAligner.result_from_record(rec)
            if exp is None:
                assert got is None
            else:
                assert got[0] == exp[0] and (got[1].fields() if got[1] else None) == exp[1] and \
                    (got[2].fields() if got[2] else None) == exp[2], (kw, r1, r2)


def test_edge_cases():
    from atropos_b200 import _abi
    from atropos_b200.adapters import Adapter, BACK
    from atropos_b200.align import Aligner, InsertAligner
    ad = Adapter(T1, BACK)
    assert len(ad.match_to_batch([])) == 0                                   # empty batch
    res = ad.match_to_batch(["", "A", T1, "", "ACGT" * 10 + T1[:7]])          # empty and ragged reads
    assert [int(s) for s in res["status"]] == [0, 0, 1, 0, 1]
    assert _tup(res[2]) == (0, 34, 0, 34, 34, 0) and _tup(res[4]) == (0, 7, 40, 47, 7, 0)
    rng = np.random.default_rng(5)
    big = fuzzgen.rand_seq(rng, 32000) + T1                                  # maximum read size class
    assert Aligner(T1, 0.1, BACK).locate(big) == oracle.locate(T1, big, 0.1, oracle.BACK)
    a64 = fuzzgen.rand_seq(rng, 64)
    r = fuzzgen.rand_seq(rng, 50) + fuzzgen.mutate(rng, a64, 0.05, 0.02, 0.02) + fuzzgen.rand_seq(rng, 20)
    assert Aligner(a64, 0.2, 15).locate(r) == oracle.locate(a64, r, 0.2, 15)
    assert InsertAligner(T1, T2).match_insert("", "") is None
    with pytest.raises(KeyError):
        InsertAligner(T1, T2).match_insert("ACGT", "AC.T")
    with pytest.raises(UnicodeEncodeError):
        Aligner(T1, 0.1).locate("ACGé")


def test_full_size_properties():
    """BASELINE's full size (10 M reads) through size-independent properties: (1) determinism / independence
    of reads: aligning a permuted batch gives the permuted results; (2) idempotence of trimming: after cutting
    at rstart the adapter found in round 1 is gone or strictly shorter; (3) planted exact adapters are reported
    at their planted position; (4) all 10 M records equal the oracle's."""
    import torch
    from atropos_b200 import _abi, synth
    from atropos_b200.adapters import Adapter, BACK
    n, L = 10_000_000, 150
    reads = synth.synth_se(n, L, seed=synth.seed_for(2), device="cuda").cpu().numpy()
    offs = np.arange(n + 1, dtype=np.int64) * L
    ad = Adapter(T1, BACK, 0.1, 3)
    res = ad.match_to_batch((reads.reshape(-1), offs))
    hit = res["status"] == _abi.ATR_ST_MATCH
    assert 0.35 < hit.mean() < 0.45
    rng = np.random.default_rng(1)
    # (4) EVERY read of the batch against the oracle (the C restatement on all host threads: a few seconds per 10 M reads)
    import os
    exp = oracle.locate_batch(T1, reads.reshape(-1), offs, 0.1, oracle.BACK, False, False, 3, 1, threads=os.cpu_count() or 1)
    assert _check_locate_array(res, exp) == int(hit.sum())
    del exp
    perm = rng.permutation(2_000_000)
    res_p = ad.match_to_batch((reads[:2_000_000][perm].reshape(-1), offs[:2_000_001]))
    assert np.array_equal(res_p, res[:2_000_000][perm])
    # (2) windows [0, rstart): a second round never finds the same or a longer adapter stretch at the same place
    win = np.zeros((2_000_000, 2), dtype=np.uint16)
    win[:, 1] = np.where(hit[:2_000_000], res["rstart"][:2_000_000], L)
    res2 = ad.match_to_batch((reads[:2_000_000].reshape(-1), offs[:2_000_001]), win=win)
    again = (res2["status"] == _abi.ATR_ST_MATCH) & hit[:2_000_000]
    assert np.all(res2["rstop"][again] <= res["rstart"][:2_000_000][again])
    # (3) planted exact adapter
    planted = reads[:1_000_000].copy()
    pos = rng.integers(0, L - 34, size=len(planted))
    a = np.frombuffer(T1.encode(), dtype=np.uint8)
    cols = pos[:, None] + np.arange(34)[None, :]
    np.put_along_axis(planted, cols, np.broadcast_to(a, cols.shape), axis=1)
    res3 = ad.match_to_batch((planted.reshape(-1), offs[:1_000_001]))
    assert np.all(res3["status"] == _abi.ATR_ST_MATCH)
    full = (res3["errors"] == 0) & (res3["matches"] == 34)
    assert np.all(res3["rstart"][full] <= pos[full])        # leftmost exact occurrence (an earlier one can exist by design)
    assert full.mean() > 0.999


def test_insert_adapter_cutter_matching_half():
    """InsertAdapterCutter.__call__ up to the point where Match objects exist (modifiers.py:391-406): insert match
    first, per-read adapter match (insert-mode adapters: max_rmp 1e-6, overlap 1, indel cost 3; trim/cli.py:667-684)
    only for the pairs without one."""
    from atropos_b200 import _abi, synth
    from atropos_b200.adapters import Adapter, BACK
    from atropos_b200.align import InsertAligner
    from atropos_b200.modifiers import InsertAdapterCutter
    from atropos_b200.util import RandomMatchProbability
    n, L = 6000, 150
    r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(3, 7), device="cpu", sub=0.03)
    reads1 = [bytes(x).decode() for x in r1.numpy()]
    reads2 = [bytes(x).decode() for x in r2.numpy()]
    reads1[5], reads2[5] = "", "ACGT"                      # shorter than min_insert_overlap -> skipped
    rmp = RandomMatchProbability()
    kw = dict(max_error_rate=0.1, min_overlap=1, indel_cost=3, match_probability=rmp, max_rmp=1e-6)
    a1, a2 = Adapter(T1, BACK, **kw), Adapter(T2, BACK, **kw)
    ikw = dict(max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1)
    cutter = InsertAdapterCutter(a1, a2, InsertAligner(T1, T2, **ikw))
    ins, fb1, fb2, need = cutter.match_batch(reads1, reads2)
    o_rmp = oracle.RandomMatchProbability()
    okw = dict(max_error_rate=0.1, min_overlap=1, indel_cost=3, match_probability=o_rmp, max_rmp=1e-6)
    o1, o2 = oracle.OracleAdapter(T1, oracle.BACK, **okw), oracle.OracleAdapter(T2, oracle.BACK, **okw)
    oia = oracle.OracleInsertAligner(T1, T2, **ikw)
    n_fb = n_ins = 0
    for i in range(n):
        if len(reads1[i]) < 1 or len(reads2[i]) < 1:
            assert ins[i]["insert"]["status"] == _abi.ATR_ST_NONE and not need[i]
            continue
        exp = oia.match_insert(reads1[i], reads2[i])
        got = InsertAligner.result_from_record(ins[i])
        if exp is not None:
            n_ins += 1
            assert got is not None and got[0] == exp[0] and not need[i]
            assert fb1[i]["status"] == _abi.ATR_ST_NONE and fb2[i]["status"] == _abi.ATR_ST_NONE
        else:
            assert got is None and need[i]
            for rec, orc_ad, read in ((fb1[i], o1, reads1[i]), (fb2[i], o2, reads2[i])):
                e = orc_ad.match_to(read)
                g = None if rec["status"] == _abi.ATR_ST_NONE else _tup(rec)
                assert g == (None if e is None else e[:6]), (i, read)
                n_fb += e is not None
    assert n_ins > 1500 and n_fb > 20


_PE_FULL = None


def _pe_oracle_range(bounds):
    """worker of test_full_size_properties_pe: OracleInsertAligner.match_insert for pairs [lo, hi) of the inherited batch"""
    lo, hi = bounds
    r1, r2, kw = _PE_FULL
    orc = oracle.OracleInsertAligner(T1, T2, **kw)
    out = np.zeros((hi - lo, 21), dtype=np.int64)
    for i in range(lo, hi):
        e = orc.match_insert(r1[i].tobytes().decode(), r2[i].tobytes().decode())
        if e is None:
            continue
        row = out[i - lo]
        row[0] = 1
        row[1:7] = e[0]
        if e[1] is not None:
            row[7] = 1
            row[8:14] = e[1]
        if e[2] is not None:
            row[14] = 1
            row[15:21] = e[2]
    return out


def _pe_check_all(res, r1, r2, kw):
    """every INSERT_DTYPE record of `res` against OracleInsertAligner.match_insert, one forked worker per host core"""
    import multiprocessing as mp
    import os
    from atropos_b200 import _abi
    global _PE_FULL
    n = len(res)
    _PE_FULL = (r1, r2, kw)
    workers = max(1, os.cpu_count() or 1)
    bounds = np.linspace(0, n, 8 * workers + 1).astype(np.int64)
    with mp.get_context("fork").Pool(workers) as pool:
        exp = np.concatenate(pool.map(_pe_oracle_range, [(int(bounds[t]), int(bounds[t + 1])) for t in range(8 * workers)]))
    _PE_FULL = None                                                # exp: [n, 21] found | insert 6 | has1, match1 6 | has2, match2 6
    found = exp[:, 0] == 1
    assert np.array_equal(res["insert"]["status"] == _abi.ATR_ST_MATCH, found)
    for c, f in enumerate(FIELDS):
        assert np.array_equal(res["insert"][f][found].astype(np.int64), exp[found, 1 + c]), ("insert", f)
    for side, o in (("match1", 7), ("match2", 14)):
        has = found & (exp[:, o] == 1)
        assert np.array_equal((res[side]["status"] == _abi.ATR_ST_MATCH) & found, has), side
        for c, f in enumerate(FIELDS):
            assert np.array_equal(res[side][f][has].astype(np.int64), exp[has, o + 1 + c]), (side, f)
    return int(found.sum())


def test_full_size_properties_pe():
    """BASELINE config 3 at full size (10 M pairs, 2 x 150): (1) all 10 M results equal the oracle's (Python restatement on
    every host core, about a minute); (2) pairs are independent: a permuted batch gives the permuted records;
    (3) swapping mates with swapped adapters mirrors the result (the overlap is symmetric under reverse
    complement: same insert size, match1 <-> match2); (4) planted error-free fragments shorter than the read are
    all found with insert size == fragment length."""
    import torch
    from atropos_b200 import _abi, synth
    from atropos_b200.align import InsertAligner
    n, L = 10_000_000, 150
    r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(3), device="cuda")
    r1, r2 = r1.cpu().numpy(), r2.cpu().numpy()
    offs = np.arange(n + 1, dtype=np.int64) * L
    kw = dict(max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1)
    ia = InsertAligner(T1, T2, **kw)
    res = ia.match_insert_batch((r1.reshape(-1), offs), (r2.reshape(-1), offs))
    st = res["insert"]["status"]
    assert set(np.unique(st)) <= {_abi.ATR_ST_NONE, _abi.ATR_ST_MATCH}
    assert 0.35 < (st == _abi.ATR_ST_MATCH).mean() < 0.45
    rng = np.random.default_rng(3)
    # (1) EVERY pair of the batch against the oracle, on all host cores (forked workers: they only run the CPU oracle)
    assert _pe_check_all(res, r1, r2, kw) == int((st == _abi.ATR_ST_MATCH).sum())
    m = 2_000_000
    perm = rng.permutation(m)
    res_p = ia.match_insert_batch((r1[:m][perm].reshape(-1), offs[:m + 1]), (r2[:m][perm].reshape(-1), offs[:m + 1]))
    assert np.array_equal(res_p, res[:m][perm])
    # (3) mirror: Hamming(rc(r2)[m-j:], r1[:j]) == Hamming(rc(r1)[m-j:], r2[:j]) for every j, and the decision
    # procedure is symmetric in (read1, adapter1) <-> (read2, adapter2)
    ia_sw = InsertAligner(T2, T1, **kw)
    res_sw = ia_sw.match_insert_batch((r2[:m].reshape(-1), offs[:m + 1]), (r1[:m].reshape(-1), offs[:m + 1]))
    assert np.array_equal(res[:m]["insert"], res_sw["insert"])
    assert np.array_equal(res[:m]["match1"][list(FIELDS)], res_sw["match2"][list(FIELDS)])
    assert np.array_equal(res[:m]["match2"][list(FIELDS)], res_sw["match1"][list(FIELDS)])
    # (4) planted fragments
    k = 500_000
    frag = rng.integers(40, L - 12, size=k)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    base = np.frombuffer(b"ACGT", dtype=np.uint8)
    F = base[rng.integers(0, 4, size=(k, L))]
    p1 = base[rng.integers(0, 4, size=(k, L))]
    p2 = base[rng.integers(0, 4, size=(k, L))]
    a1 = np.frombuffer(T1.encode(), dtype=np.uint8)
    a2 = np.frombuffer(T2.encode(), dtype=np.uint8)
    col = np.arange(L)[None, :]
    rel = col - frag[:, None]
    p1 = np.where(rel < 0, F, np.where(rel < len(a1), a1[np.clip(rel, 0, len(a1) - 1)], p1))
    Frc = comp[np.take_along_axis(F, np.clip(frag[:, None] - 1 - col, 0, L - 1), axis=1)]
    p2 = np.where(rel < 0, Frc, np.where(rel < len(a2), a2[np.clip(rel, 0, len(a2) - 1)], p2))
    res4 = ia.match_insert_batch((np.ascontiguousarray(p1).reshape(-1), offs[:k + 1]),
                                 (np.ascontiguousarray(p2).reshape(-1), offs[:k + 1]))
    assert np.all(res4["insert"]["status"] == _abi.ATR_ST_MATCH)
    ok = res4["match1"]["rstart"] == frag
    assert ok.mean() > 0.999        # a chance longer overlap of the random tails can win by probability
    assert np.all(res4["match1"]["status"][ok] == _abi.ATR_ST_MATCH) and np.all(res4["match2"]["rstart"][ok] == frag[ok])


def test_batched_trim_consumes_gpu_records():
    """f-1/f-3: trimming windows and adapter statistics from the GPU's records equal those from the oracle's"""
    from atropos_b200 import trim
    from atropos_b200.adapters import Adapter
    from atropos_b200.modifiers import AdapterCutter
    from test_trim_vs_reference import SPECS, _oracle_rounds
    rng = np.random.default_rng(2025)
    reads = []
    for _ in range(3000):
        s, w = SPECS[int(rng.integers(0, len(SPECS)))]
        body = fuzzgen.read_with_adapter(rng, s, int(rng.integers(0, 160)), n_rate=0.01)
        if w != oracle.BACK and rng.random() < 0.6:
            body = (fuzzgen.mutate(rng, s, 0.03, 0.01, 0.01) + body)[:150]
        reads.append(body)
    gpu_adapters = [Adapter(s, w) for s, w in SPECS]
    cutter = AdapterCutter(gpu_adapters, times=2)
    blob = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=offs[1:])
    g_rounds = cutter.match_rounds_batch((blob, offs))
    o_rounds = _oracle_rounds(reads, [oracle.OracleAdapter(s, w) for s, w in SPECS], 2)
    ff = [a._front_flag for a in gpu_adapters]
    g = trim.apply_rounds(blob, offs, g_rounds, ff)
    o = trim.apply_rounds(blob, offs, o_rounds, ff)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and g[3] == o[3]
    for a, b in zip(g[2], o[2]):
        assert a.lengths_front == b.lengths_front and a.lengths_back == b.lengths_back
        assert a.errors_front == b.errors_front and a.errors_back == b.errors_back
        assert a.adjacent_bases == b.adjacent_bases
    assert g[3] > 1000


def test_packed_host_entry_point_equals_ascii_entry_point():
    """atr_locate_batch_host_packed (reads packed on the host by atr_pack_reads_host; only codes / index cross PCIe,
    escaped reads' bytes ride along) == atr_locate_batch_host: single adapter and panel, windows, ragged and
    fixed-length batches, lower case / other bytes, with and without the ASCII for the escaped reads."""
    from atropos_b200 import _abi, engine
    from atropos_b200.adapters import Adapter, BACK, FRONT
    from atropos_b200.modifiers import AdapterCutter
    rng = np.random.default_rng(77)
    for fixed in (True, False):
        reads = []
        for i in range(30000):
            L = 150 if fixed else int(rng.integers(0, 260))
            r = fuzzgen.read_with_adapter(rng, T1, L, n_rate=0.01)
            if not fixed and rng.random() < 0.03:
                r = r.lower() if rng.random() < 0.5 else r.replace("A", ".", 1)
            reads.append(r)
        ascii, offsets = engine.encode_reads(reads)
        for adapters, fold in (([Adapter(T1, BACK)], True), ([Adapter(T1, BACK), Adapter("TGGAATTCTCGGGTGCCAAGG", BACK),
                                                              Adapter("AATGATACGGCGACCACCGA", FRONT)], False)):
            aset = AdapterCutter(adapters)._adapterset()
            win = None
            if not fixed:
                lens_ = np.diff(offsets)
                lo = (rng.random(len(reads)) * lens_ * 0.3).astype(np.uint16)
                win = np.stack([lo, lens_.astype(np.uint16)], axis=1)
            exp = aset.locate_host(ascii, offsets, win=win, fold_case=fold)
            codes, woff, lens = engine.pack_reads_host(ascii, offsets, fold_case=fold)
            got = aset.locate_host_packed(codes, woff, lens, win=win, ascii=ascii, offsets=offsets, fold_case=fold)
            assert np.array_equal(got, exp)
            got2 = aset.locate_host_packed(codes, woff, lens, win=win, fold_case=fold)
            esc = (lens & 0x8000) != 0
            assert np.array_equal(got2[~esc], exp[~esc])
            if esc.any():
                assert (got2["status"][esc] == _abi.ATR_ST_ESCAPED).all()


def test_panel_cfg4_all_eight_adapters():
    """BASELINE cfg 4's panel exactly as bench.py runs it (three 3' adapters incl. the 58-mer, two anchored 5', one
    unanchored 5', the duplicate and the 13-mer): best match per read on the GPU == the reference's rule
    (modifiers.py:107-122: strictly more matches wins, the first adapter on ties) over the oracle's match_to."""
    import bench
    from atropos_b200 import _abi, adapters as ad_mod, engine
    from atropos_b200.modifiers import AdapterCutter
    rng = np.random.default_rng(404)
    reads = []
    for i in range(12000):
        which = int(rng.integers(0, len(bench.PANEL)))
        seq, where = bench.PANEL[which]
        body = fuzzgen.read_with_adapter(rng, seq, 150, n_rate=0.005) if where == "BACK" else fuzzgen.rand_seq(rng, 150)
        r = rng.random()
        if where != "BACK" and r < 0.7:
            body = (fuzzgen.mutate(rng, seq, 0.03, 0.01, 0.01) + body)[:150]
        elif r < 0.1:
            body = (bench.PANEL[3][0] + body)[:150]
        reads.append(body)
    cutter = AdapterCutter([ad_mod.Adapter(s, getattr(ad_mod, w), max_error_rate=0.1, min_overlap=3) for s, w in bench.PANEL])
    got = cutter.best_match_batch(reads)
    oads = [oracle.OracleAdapter(s, getattr(oracle, w), 0.1, 3) for s, w in bench.PANEL]
    hits = 0
    winners = set()
    for i, seq in enumerate(reads):
        best, bi = None, -1
        for ai, oa in enumerate(oads):
            m = oa.match_to(seq)
            if m is not None and (best is None or m[4] > best[4]):
                best, bi = m, ai
        g = got[i]
        if best is None:
            assert int(g["status"]) == _abi.ATR_ST_NONE, (i, seq)
        else:
            hits += 1
            winners.add(bi)
            assert int(g["status"]) == _abi.ATR_ST_MATCH and _tup(g) == tuple(best[:6]) and int(g["adapter"]) == bi, (i, seq, best, bi)
    assert hits > 6000 and len(winners) >= 6


_PANEL_FULL = None


def _panel_oracle_range(bounds):
    """worker of test_panel_cfg4_two_million_reads: the reference's best-match rule over the oracle's match_to"""
    import bench
    lo, hi = bounds
    reads = _PANEL_FULL
    oads = [oracle.OracleAdapter(s, getattr(oracle, w), 0.1, 3) for s, w in bench.PANEL]
    out = np.zeros((hi - lo, 8), dtype=np.int64)
    for i in range(lo, hi):
        seq = reads[i].tobytes().decode()
        best, bi = None, -1
        for ai, oa in enumerate(oads):
            m = oa.match_to(seq)
            if m is not None and (best is None or m[4] > best[4]):      # modifiers.py:116-121
                best, bi = m, ai
        if best is not None:
            out[i - lo, 0] = 1
            out[i - lo, 1:7] = best[:6]
            out[i - lo, 7] = bi
    return out


def test_panel_cfg4_two_million_reads():
    """BASELINE cfg 4's data and panel as bench.py builds them (40 % of the reads run into the TruSeq adapter, 10 % start
    with one of the two anchored 5' adapters): the first 2 M reads of the shard, every record and winning adapter against
    the oracle (forked workers, one per host core)."""
    import multiprocessing as mp
    import os
    import bench
    import torch
    from atropos_b200 import _abi, adapters as ad_mod, synth
    from atropos_b200.modifiers import AdapterCutter
    n, L = 2_000_000, 150
    reads = synth.synth_se(n, L, bench.ADAPTER, seed=synth.seed_for(4, 0), device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(synth.seed_for(4, 0) + 7)
    pick = torch.rand(n, generator=g, device="cuda")
    for k, seq in enumerate([bench.PANEL[3][0], bench.PANEL[4][0]]):
        rows = torch.nonzero((pick >= 0.05 * k) & (pick < 0.05 * (k + 1))).squeeze(1)
        reads[rows, :len(seq)] = torch.tensor(list(seq.encode()), dtype=torch.uint8, device="cuda")[None, :]
    reads = reads.cpu().numpy()
    cutter = AdapterCutter([ad_mod.Adapter(s, getattr(ad_mod, w), max_error_rate=0.1, min_overlap=3) for s, w in bench.PANEL])
    got = cutter.best_match_batch((reads.reshape(-1), np.arange(n + 1, dtype=np.int64) * L))
    global _PANEL_FULL
    _PANEL_FULL = reads
    workers = max(1, os.cpu_count() or 1)
    bounds = np.linspace(0, n, 8 * workers + 1).astype(np.int64)
    with mp.get_context("fork").Pool(workers) as pool:
        exp = np.concatenate(pool.map(_panel_oracle_range, [(int(bounds[t]), int(bounds[t + 1])) for t in range(8 * workers)]))
    _PANEL_FULL = None
    found = exp[:, 0] == 1
    assert np.array_equal(got["status"] == _abi.ATR_ST_MATCH, found)
    assert np.array_equal(got["status"] == _abi.ATR_ST_NONE, ~found)
    for c, f in enumerate(FIELDS):
        assert np.array_equal(got[f][found].astype(np.int64), exp[found, 1 + c]), f
    assert np.array_equal(got["adapter"][found].astype(np.int64), exp[found, 7])
    winners = np.bincount(exp[found, 7], minlength=len(bench.PANEL))
    assert found.mean() > 0.4 and (winners > 0).sum() >= 4, winners
