"""The per-read / per-pair device functions (atropos_b200/csrc/*_core.cuh are __host__ __device__), compiled
for the CPU by tests/host_sim and compared with the oracle. This is how kernel LOGIC is checked in the
GPU-less build container; the same comparisons run through the real kernels in test_gpu_parity.py."""
import numpy as np
import pytest

import fuzzgen
import golden_util
import hostsim
from atropos_b200 import _abi
from atropos_b200.adapters import Adapter, ANYWHERE, BACK, FRONT, PREFIX, SUFFIX
from atropos_b200.align import InsertAligner
from atropos_b200.util import RandomMatchProbability
from oracle import oracle

T1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"
T2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT"


def _tup(rec):
    return tuple(int(rec[k]) for k in ("astart", "astop", "rstart", "rstop", "matches", "errors"))


@pytest.mark.parametrize("route", [0, 1])
def test_locate_fuzz(route):
    used = bad = 0
    for c in fuzzgen.locate_cases(21, 6000):
        exp = oracle.locate(c["reference"], c["query"], c["max_error_rate"], c["flags"], False, False,
                            c["min_overlap"], c["indel_cost"])
        d, keep = _abi.make_adapter_desc(c["reference"], c["max_error_rate"], c["flags"], False, False,
                                         c["min_overlap"], c["indel_cost"])
        got, k1a, _ = hostsim.locate(c["query"], d, route=route)
        used += k1a
        bad += got != exp
    assert bad == 0
    assert (used > 5000) if route == 0 else used == 0


@pytest.mark.parametrize("route", [0, 1])
def test_locate_wildcard_fuzz(route):
    bad = 0
    for c in fuzzgen.wildcard_locate_cases(22, 4000):
        exp = oracle.locate(c["reference"], c["query"], c["max_error_rate"], c["flags"], c["wildcard_ref"],
                            c["wildcard_query"], c["min_overlap"], c["indel_cost"])
        d, keep = _abi.make_adapter_desc(c["reference"], c["max_error_rate"], c["flags"], c["wildcard_ref"],
                                         c["wildcard_query"], c["min_overlap"], c["indel_cost"])
        got, _, _ = hostsim.locate(c["query"], d, route=route)
        bad += got != exp
    assert bad == 0


def test_locate_golden():
    for c in golden_util.load("locate"):
        d, keep = _abi.make_adapter_desc(c["reference"], c["max_error_rate"], c["flags"], c["wildcard_ref"],
                                         c["wildcard_query"], c["min_overlap"], c["indel_cost"])
        got, _, _ = hostsim.locate(c["query"], d)
        assert got == (None if c["expect"] is None else tuple(c["expect"])), c


def test_locate_escaped_and_long_reads():
    """bytes the 4-bit code cannot hold (lower case, U, '.') and reads > 4000 nt take the byte-exact path"""
    rng = np.random.default_rng(23)
    d, keep = _abi.make_adapter_desc(T1, 0.1, BACK, False, False, 3, 1)
    for _ in range(300):
        read = list(fuzzgen.read_with_adapter(rng, T1, 120))
        for i in range(len(read)):
            if rng.random() < 0.05:
                read[i] = "acgtnU.x"[int(rng.integers(0, 8))]
        read = "".join(read)
        got, k1a, _ = hostsim.locate(read, d)
        assert got == oracle.locate(T1, read, 0.1, BACK, False, False, 3, 1)
    long_read = fuzzgen.rand_seq(rng, 4500) + T1[:20] + fuzzgen.rand_seq(rng, 100)
    got, k1a, _ = hostsim.locate(long_read, d)
    assert not k1a and got == oracle.locate(T1, long_read, 0.1, BACK, False, False, 3, 1)
    ok_read = long_read[-3900:]
    got, k1a, _ = hostsim.locate(ok_read, d)
    assert k1a and got == oracle.locate(T1, ok_read, 0.1, BACK, False, False, 3, 1)


def test_locate_long_adapter_and_high_rate():
    rng = np.random.default_rng(24)
    for m in (64, 65, 100, 300):
        ad = fuzzgen.rand_seq(rng, m)
        for rate in (0.1, 0.2, 1.0, 2.5):
            d, keep = _abi.make_adapter_desc(ad, rate, 15, False, False, 3, 1)
            for _ in range(6):
                read = fuzzgen.read_with_adapter(rng, ad, int(rng.integers(50, 400)))
                got, k1a, _ = hostsim.locate(read, d)
                assert got == oracle.locate(ad, read, rate, 15, False, False, 3, 1), (m, rate, read)
                assert bool(k1a) == (m <= 64 and int(rate * m) <= 126)


def test_windows():
    """a (lo, hi) window must behave exactly like slicing the read (times > 1 rounds, linked adapters)"""
    rng = np.random.default_rng(25)
    for _ in range(400):
        where = [BACK, FRONT, ANYWHERE, PREFIX, SUFFIX][int(rng.integers(0, 5))]
        ad = fuzzgen.rand_seq(rng, int(rng.integers(5, 40)))
        read = fuzzgen.read_with_adapter(rng, ad, 100)
        lo = int(rng.integers(0, 60))
        hi = int(rng.integers(lo, 101))
        d, keep = _abi.make_adapter_desc(ad, 0.15, where, False, False, 3, 1)
        got, _, _ = hostsim.locate(read, d, lo=lo, hi=hi)
        assert got == oracle.locate(ad, read[lo:hi], 0.15, where, False, False, 3, 1)


@pytest.mark.parametrize("where", [BACK, FRONT, ANYWHERE, PREFIX, SUFFIX])
@pytest.mark.parametrize("route", [0, 1])
def test_match_to_semantics(where, route):
    rng = np.random.default_rng(30 + where)
    rmp, rmp_o = RandomMatchProbability(), oracle.RandomMatchProbability()
    found = 0
    for _ in range(250):
        wild = rng.random() < 0.3
        seq = fuzzgen.rand_seq(rng, int(rng.integers(3, 50)), "ACGTACGTACGTNRY" if wild else "ACGT")
        kw = dict(max_error_rate=float(rng.choice([0.0, 0.1, 0.12, 0.2])), min_overlap=int(rng.choice([1, 3, 5])),
                  read_wildcards=bool(rng.random() < 0.3), adapter_wildcards=bool(rng.random() < 0.7),
                  indels=bool(rng.random() < 0.7), indel_cost=int(rng.choice([1, 1, 3])))
        max_rmp = [None, 1e-6, 1e-3][int(rng.integers(0, 3))]
        mine = Adapter(seq, where, match_probability=rmp, max_rmp=max_rmp, **kw)
        orc = oracle.OracleAdapter(seq, where, match_probability=rmp_o, max_rmp=max_rmp, **kw)
        d, keep = mine.descriptor()
        proj = "".join(ch if ch in "ACGT" else "ACGT"[int(rng.integers(0, 4))] for ch in seq)
        for _ in range(6):
            read = fuzzgen.read_with_adapter(rng, proj if rng.random() < 0.7 else seq, int(rng.integers(0, 160)),
                                             n_rate=0.02)
            if rng.random() < 0.15:
                read = read.lower()
            exp = orc.match_to(read)
            got, _, _ = hostsim.locate(read, d, route=route, fold_case=True)
            assert got == (None if exp is None else exp[:6]), (seq, where, kw, max_rmp, read)
            found += exp is not None
    assert found > 20


def test_match_to_golden():
    rmp = RandomMatchProbability()
    for c in golden_util.load("match_to"):
        d, keep = Adapter(c["sequence"], c["where"], match_probability=rmp, max_rmp=c["max_rmp"], **c["kw"]).descriptor()
        for read, exp in zip(c["reads"], c["expect"]):
            got, _, _ = hostsim.locate(read, d, fold_case=True)
            assert got == (None if exp is None else tuple(exp[:6])), (c, read)


def test_best_match_reduction():
    """AdapterCutter._best_match over a panel: run adapters in order with reduce=1 on the same record"""
    rng = np.random.default_rng(40)
    seqs = [T1, T2, "TGGAATTCTCGGGTGCCAAGG", "AGATCGGAAGAGC"]
    mine = [Adapter(s, BACK) for s in seqs]
    orc = [oracle.OracleAdapter(s, oracle.BACK) for s in seqs]
    descs = [a.descriptor() for a in mine]
    for _ in range(1500):
        read = fuzzgen.read_with_adapter(rng, seqs[int(rng.integers(0, 4))], 100)
        prev = None
        for i, (d, keep) in enumerate(descs):
            _, _, prev = hostsim.locate(read, d, fold_case=True, prev=prev, adapter_index=i)
        exp = oracle.best_match(orc, read)
        if exp is None:
            assert prev.status == _abi.ATR_ST_NONE
        else:
            assert prev.adapter == exp[0] and hostsim.decode(prev) == exp[1][:6]


@pytest.mark.parametrize("cfg", range(5))
def test_match_insert(cfg):
    kw = [dict(max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1), dict(),
          dict(max_insert_mismatch_frac=0.3, max_adapter_mismatch_frac=0.2, min_insert_overlap=10, adapter_wildcards=False),
          dict(read_wildcards=True), dict(read_wildcards=True, adapter_wildcards=False)][cfg]
    d, keep = InsertAligner(T1, T2, **kw).descriptor(160)
    orc = oracle.OracleInsertAligner(T1, T2, **kw)
    matched = packed = 0
    rng = np.random.default_rng(50 + cfg)
    for r1, r2 in fuzzgen.insert_pairs(500 + cfg, 700, T1, T2, err=[0.01, 0.03, 0.08, 0.03, 0.03][cfg]):
        if rng.random() < 0.05:        # bytes outside the packed alphabet: lower case, X (KeyError), '.'
            pos = int(rng.integers(0, len(r2)))
            r2 = r2[:pos] + "aX.n"[int(rng.integers(0, 4))] + r2[pos + 1:]
        try:
            exp = orc.match_insert(r1, r2)
        except KeyError:
            exp = "KEYERROR"
        for route in (0, 1):
            rec, used = hostsim.match_insert(d, r1, r2, route)
            packed += used
            st = int(rec["insert"]["status"])
            if exp == "KEYERROR":
                assert st == _abi.ATR_ST_KEYERROR
            elif exp is None:
                assert st == _abi.ATR_ST_NONE
            else:
                matched += 1
                assert st == _abi.ATR_ST_MATCH and _tup(rec["insert"]) == exp[0]
                for e, g in ((exp[1], rec["match1"]), (exp[2], rec["match2"])):
                    if e is None:
                        assert int(g["status"]) == _abi.ATR_ST_NONE
                    else:
                        assert int(g["status"]) == _abi.ATR_ST_MATCH and _tup(g) == e
    assert matched > 200 and packed > 500


def test_match_insert_golden():
    for g in golden_util.load("match_insert"):
        d, keep = InsertAligner(g["adapter1"], g["adapter2"], **g["kw"]).descriptor(160)
        for (r1, r2), exp in zip(g["pairs"], g["expect"]):
            rec, _ = hostsim.match_insert(d, r1, r2)
            if exp is None:
                assert int(rec["insert"]["status"]) == _abi.ATR_ST_NONE
                continue
            assert list(_tup(rec["insert"])) == exp[0]
            for e, gm in ((exp[1], rec["match1"]), (exp[2], rec["match2"])):
                if e is None:
                    assert int(gm["status"]) == _abi.ATR_ST_NONE
                else:
                    assert list(_tup(gm)) == e[:6]


def test_multi_locate_general_flags():
    rng = np.random.default_rng(60)
    for _ in range(3000):
        alpha = rng.choice(["ACGT", "AC", "A"])
        m = int(rng.integers(1, 80))
        n = m if rng.random() < 0.6 else int(rng.integers(1, 80))
        ref, q = fuzzgen.rand_seq(rng, m, alpha), fuzzgen.rand_seq(rng, n, alpha)
        if rng.random() < 0.6:
            ov = int(rng.integers(1, min(m, n) + 1))
            q = (fuzzgen.mutate(rng, ref[m - ov:], 0.05, 0, 0, alpha) + q)[:n]
        rate = float(rng.choice([0, 0.1, 0.2, 0.3]))
        flags = int(rng.choice([9, 9, 15, 14, 11, 8, 2]))
        mo, mm = int(rng.choice([1, 3, 10])), int(rng.choice([100, 5, 1]))
        assert oracle.multi_locate(ref, q, rate, flags, mo, mm) == hostsim.multi_locate(ref, q, rate, flags, mo, mm)
    for c in golden_util.load("multi_locate"):
        got = hostsim.multi_locate(c["reference"], c["query"], c["max_error_rate"], c["flags"], c["min_overlap"])
        assert got == (None if c["expect"] is None else [tuple(t) for t in c["expect"]])


@pytest.mark.parametrize("where", [PREFIX, SUFFIX])
def test_anchored_piece_filter(where):
    """PREFIX / SUFFIX adapters with indels take the fixed-position piece filter (k_filter_anchor) and the register
    DP only for its survivors: same answers as the oracle, and the filter really rejects most adapter-free reads."""
    rng = np.random.default_rng(400 + where)
    took = passed = found = 0
    for _ in range(400):
        m = int(rng.integers(8, 60))
        seq = fuzzgen.rand_seq(rng, m, "ACGT")
        rate = float(rng.choice([0.0, 0.1, 0.15, 0.2]))
        ic = int(rng.choice([1, 1, 3]))
        mo = int(rng.choice([1, 3, 5]))
        d, keep = _abi.make_adapter_desc(seq, rate, where, False, False, mo, ic)
        for _ in range(8):
            n = int(rng.integers(0, 160))
            body = fuzzgen.rand_seq(rng, n, "ACGT")
            r = rng.random()
            if r < 0.6:                       # the adapter (mutated, possibly shifted by an indel) at the anchored end
                mut = fuzzgen.mutate(rng, seq, sub=rate / 2, ins=rate / 4, dele=rate / 4)
                body = (mut + body)[:max(n, len(mut))] if where == PREFIX else (body + mut)[-max(n, len(mut)):]
            elif r < 0.7:
                body = (seq[:m // 2] + body) if where == PREFIX else (body + seq[m // 2:])
            exp = oracle.locate(seq, body, rate, where, False, False, mo, ic)
            got, used, _ = hostsim.locate(body, d)
            assert got == exp, (seq, rate, ic, mo, body)
            took += used >= 20
            passed += used == 21
            found += exp is not None
    # SUFFIX with unit indel cost is taken by the funnel (fused_ok); only its indel_cost 3 third comes here
    assert took > (2000 if where == PREFIX else 800) and found > 500
    assert passed < 0.85 * took           # the filter is selective


def test_match_insert_long_reads_high_rate():
    """BASELINE config 5 shape (2x300, rate 0.15 => bounds up to 45): the overlaps whose bound a 32-base look cannot
    exceed are finished inline instead of being parked (ATR_K2_INLINE_THR); candidates must still come out in order"""
    from atropos_b200 import synth
    kw = dict(max_insert_mismatch_frac=0.15, max_adapter_mismatch_frac=0.15)
    L = 300
    d, keep = InsertAligner(T1, T2, **kw).descriptor(L)
    orc = oracle.OracleInsertAligner(T1, T2, **kw)
    r1, r2 = synth.synth_pe(260, L, seed=77, device="cpu", sub=0.03)
    r1, r2 = r1.numpy(), r2.numpy()
    rng = np.random.default_rng(78)
    matched = 0
    for i in range(260):
        a, b = bytes(r1[i]).decode(), bytes(r2[i]).decode()
        if i % 5 == 0:                   # low-complexity mates: many candidates, the 100-candidate cap
            a = ("AC" * 150)[:int(rng.integers(200, 301))]
            b = ("GT" * 150)[:int(rng.integers(200, 301))]
        elif i % 7 == 0:
            a, b = a[:int(rng.integers(120, 301))], b[:int(rng.integers(120, 301))]
        exp = orc.match_insert(a, b)
        rec, used = hostsim.match_insert(d, a, b, 0)
        assert used
        st = int(rec["insert"]["status"])
        if exp is None:
            assert st == _abi.ATR_ST_NONE, i
        else:
            matched += 1
            assert st == _abi.ATR_ST_MATCH and _tup(rec["insert"]) == exp[0], i
            for e, g in ((exp[1], rec["match1"]), (exp[2], rec["match2"])):
                assert (int(g["status"]) == _abi.ATR_ST_NONE) if e is None else (_tup(g) == e), i
    assert matched > 60


@pytest.mark.parametrize("where", [BACK, FRONT, ANYWHERE])
def test_filter_only_funnel_for_dearer_indels(where):
    """indel cost 3 (insert mode's fallback adapters) and 100000 (--no-indels): the funnel's unit-cost filter stage
    in front of the register DP -- same answers as the oracle, most adapter-free reads never reach the DP"""
    rng = np.random.default_rng(600 + where)
    took = dp = found = 0
    for _ in range(300):
        m = int(rng.integers(8, 62))
        seq = fuzzgen.rand_seq(rng, m, "ACGT")
        rate = float(rng.choice([0.0, 0.1, 0.15, 0.2]))
        ic = int(rng.choice([3, 3, 100000]))
        mo = int(rng.choice([1, 3, 5]))
        d, keep = _abi.make_adapter_desc(seq, rate, where, False, False, mo, ic)
        for _ in range(8):
            read = fuzzgen.read_with_adapter(rng, seq, int(rng.integers(0, 160)), n_rate=0.01)
            exp = oracle.locate(seq, read, rate, where, False, False, mo, ic)
            got, used, _ = hostsim.locate(read, d)
            assert got == exp, (seq, rate, ic, mo, read)
            took += used >= 30
            dp += used == 31
            found += exp is not None
    assert took > 2000 and found > 400
    assert dp < 0.9 * took


@pytest.mark.parametrize("adapter,rate,min_overlap", [
    (T1, 0.1, 3), (T1, 0.12, 1), ("AGATCGGAAGAGC", 0.1, 3), ("TGGAATTCTCGGGTGCCAAGG", 0.1, 3), (T1, 0.2, 5),
    ("ACGTACGTACGTACGTACGTAAAA", 0.13, 3), ("A" * 30, 0.1, 3), ("GATCGGAAGAGCACACGTCTGAACTCCAGTCACGATC", 0.09, 3),
    ("CTGTCTCTTATACACATCTCCGAGCCCACGAGAC", 0.1, 3),
    # pieces over more than 32 rows (the q-gram form alone: 64-bit tail pass and gate)
    (T2, 0.1, 3), (T2, 0.1, 1), (T2[:45], 0.12, 3), ("GATTACAGGCTTAACCGGTATCGATCGGAAGAGCTTGACCAGTACGGATCCTTAGGCAAGTCCA", 0.1, 5)])
def test_band_margins_fuzz(adapter, rate, min_overlap):
    """the diagonals the banded kernels keep (margin k around the piece hits, floor(i*rate) around the last-column
    candidates, hits no candidate can pass through dropped) against the oracle's full DP"""
    d, keep = _abi.make_adapter_desc(adapter, rate, 14, False, False, min_overlap, 1)
    banded = 0
    for read in fuzzgen.band_cases(4100 + len(adapter), adapter, 4000):
        exp = oracle.locate(adapter, read, rate, 14, False, False, min_overlap, 1)
        got, path, _ = hostsim.locate(read, d, route=0)
        assert got == exp, (adapter, rate, read, exp, got, path)
        banded += path in (3, 13)
    assert banded > 500


@pytest.mark.parametrize("adapter,rate,min_overlap,step", [
    (T1, 0.1, 3, 3), ("GATCGGAAGAGCACACGTCTGAACTCCA", 0.08, 1, 3), ("TGGAATTCTCGGGTGCCAAGG", 0.1, 3, 2), ("GTTCAGAGTTCTACAGTCCGACGATC", 0.1, 3, 3),
    ("ACACTCTTTCCCTACACGACGCTCTTCCGATCT", 0.1, 3, 3), ("ACGTACGTACGTACGTACGTAAAAACGTACGTAC", 0.1, 3, 3),
    ("A" * 34, 0.1, 3, 3), ("CTGTCTCTTATACACATCTCCGAGCCCACGAGAC", 0.06, 5, 3)])
def test_qgram_filter_equals_shift_and(adapter, rate, min_overlap, step):
    """the q-gram sampling form of the funnel's first stage finds exactly the automaton's piece hits: same hit range,
    same class / band / window for every read, windows included; its need-tail gate fires at least where the
    automaton's does"""
    rng = np.random.default_rng(len(adapter) * 7 + step)
    d, keep = _abi.make_adapter_desc(adapter, rate, 14, False, False, min_overlap, 1)
    seen = {0: 0, 1: 0, 2: 0, 3: 0}
    reads = list(fuzzgen.band_cases(5100 + len(adapter), adapter, 2500))
    for _ in range(1500):
        reads.append(fuzzgen.read_with_adapter(rng, adapter, int(rng.integers(0, 330)), n_rate=0.01))
    for read in reads:
        lo = hi = None
        if rng.random() < 0.25 and len(read) > 4:
            lo = int(rng.integers(0, len(read) // 2))
            hi = int(rng.integers(lo, len(read) + 1))
        res = hostsim.filter_compare(read, d, lo=lo or 0, hi=hi)
        assert res is not None
        sa, qg, sa_h, qg_h, sa_tail, qg_tail, st = res
        assert st == step
        assert sa_h == qg_h, (adapter, read, lo, hi, sa_h, qg_h)
        assert qg_tail >= sa_tail, (adapter, read, lo, hi)
        assert sa == qg, (adapter, read, lo, hi, sa, qg)
        seen[sa[0]] += 1
    assert seen[0] > 100 and seen[2] > 100 and (seen[3] > 20 or len(set(adapter)) < 4 or "ACGTACGT" in adapter), seen


@pytest.mark.parametrize("adapter,rate,min_overlap", [
    ("AATGATACGGCGACCACCGA", 0.1, 3), ("GTTCAGAGTTCTACAGTCCGACGATC", 0.1, 3), ("AGATCGGAAGAGCACACGTCTGAACTCCAGTC", 0.12, 4),
    ("ACGTACGTACGTACGTACGT", 0.1, 3), ("A" * 20, 0.1, 3), ("AATGATACGGCGACCACCGA", 0.2, 1)])
def test_front_adapters_shift_and_form(adapter, rate, min_overlap):
    """unanchored 5' adapters: the Shift-And pieces for the full-length occurrences + the exact Myers pass over the first
    m + k columns for the partial ones at the read start (front_filter) against the oracle's full DP -- adapter suffixes
    at the read start, the adapter within k columns of the start, in the middle, sticking out of the read end, twice"""
    rng = np.random.default_rng(len(adapter) * 31 + int(rate * 100))
    d, keep = _abi.make_adapter_desc(adapter, rate, oracle.FRONT, False, False, min_overlap, 1)
    through, found = 0, 0
    for _ in range(4000):
        L = int(rng.integers(1, 160))
        kind = rng.random()
        if kind < 0.4:
            suf = adapter[len(adapter) - int(rng.integers(1, len(adapter) + 1)):]
            mut = fuzzgen.mutate(rng, suf, sub=float(rng.choice([0, 0.05, 0.1, 0.2])), ins=float(rng.choice([0, 0.05])),
                                 dele=float(rng.choice([0, 0.05]))) or suf
            read = (fuzzgen.rand_seq(rng, int(rng.integers(0, 4))) if rng.random() < 0.3 else "") + mut + fuzzgen.rand_seq(rng, L)
        elif kind < 0.8:
            mut = fuzzgen.mutate(rng, adapter, sub=float(rng.choice([0, 0.03, 0.08, 0.15])), ins=float(rng.choice([0, 0.04])),
                                 dele=float(rng.choice([0, 0.04]))) or adapter
            pos = int(rng.choice([0, 1, 2, 3, 4, 5, int(rng.integers(0, 100)), L]))
            read = fuzzgen.rand_seq(rng, pos) + mut + fuzzgen.rand_seq(rng, int(rng.integers(0, 40)))
            if rng.random() < 0.3:
                read = read[:max(1, len(read) - int(rng.integers(0, len(adapter))))]
            if rng.random() < 0.2:
                read = adapter[int(rng.integers(1, len(adapter))):] + read
        else:
            read = fuzzgen.rand_seq(rng, L)
        exp = oracle.locate(adapter, read, rate, oracle.FRONT, False, False, min_overlap, 1)
        got, path, _ = hostsim.locate(read, d, route=0)
        assert got == exp, (adapter, rate, read, exp, got, path)
        through += path >= 10
        found += exp is not None
    assert found > 1000
    assert through == 4000 or rate == 0.2          # rate 0.2 on a 20-mer: 5 pieces of 4 rows, the Myers filter stays


@pytest.mark.parametrize("rate,L", [(0.25, 150), (0.5, 120), (0.9, 64), (0.2, 300), (0.05, 200)])
def test_match_insert_lookahead_budgets(rate, L):
    """the 2-bit look-ahead of the packed insert scan may never reach past the overlap (high rates: the budget of a short
    overlap asks for more look-ahead words than the overlap has) and never rejects a real candidate (N, IUPAC codes and
    low-complexity mates included): bit-exact against the oracle at error rates the benchmark does not touch"""
    from atropos_b200 import synth
    kw = dict(max_insert_mismatch_frac=rate, max_adapter_mismatch_frac=min(rate, 0.3))
    d, keep = InsertAligner(T1, T2, **kw).descriptor(L)
    orc = oracle.OracleInsertAligner(T1, T2, **kw)
    r1, r2 = synth.synth_pe(220, L, seed=int(rate * 100) + L, device="cpu", sub=min(0.3, rate * 0.6), n_rate=0.01)
    r1, r2 = r1.numpy(), r2.numpy()
    rng = np.random.default_rng(int(rate * 1000) + L)
    matched = 0
    for i in range(220):
        a, b = bytes(r1[i]).decode(), bytes(r2[i]).decode()
        if i % 6 == 0:
            a = ("ACN" * 150)[:int(rng.integers(33, L + 1))]
            b = ("GTR" * 150)[:int(rng.integers(33, L + 1))]
        elif i % 4 == 0:
            a, b = a[:int(rng.integers(1, L + 1))], b[:int(rng.integers(1, L + 1))]
        exp = orc.match_insert(a, b)
        rec, used = hostsim.match_insert(d, a, b, 0)
        assert used
        st = int(rec["insert"]["status"])
        if exp is None:
            assert st == _abi.ATR_ST_NONE, (i, a, b)
        else:
            matched += 1
            assert st == _abi.ATR_ST_MATCH and _tup(rec["insert"]) == exp[0], (i, a, b)
            for e, g in ((exp[1], rec["match1"]), (exp[2], rec["match2"])):
                assert (int(g["status"]) == _abi.ATR_ST_NONE) if e is None else (_tup(g) == e), i
    assert matched > 10
