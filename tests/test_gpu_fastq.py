"""GPU parity of the FASTQ-in -> trimmed-FASTQ-out path (atr_trim_fastq_host, rows f-1/f-2/f-3) through the C ABI:
the reference command line's golden outputs, chunk-boundary independence, and a 1 M-read run checked against
the (separately parity-tested) match records."""
import numpy as np
import pytest

import fastq_cases
from atropos_b200 import _abi, fastq, synth
from atropos_b200.adapters import BACK, Adapter

pytestmark = pytest.mark.gpu

CASES = fastq_cases.cases()


@pytest.mark.parametrize("chunk", [0, 4096, 20000])
@pytest.mark.parametrize("case", CASES, ids=[c["label"] for c in CASES])
def test_against_reference_cli(case, chunk):
    adapters = fastq_cases.adapters_of(case)
    text = case["text"].encode("latin-1")
    res = case["result"]
    tr = fastq.FastqTrimmer(adapters, times=case["times"], chunk_bytes=chunk, **case.get("read_ops", {}))
    if "error" in res:
        with pytest.raises(fastq.FormatError) as ei:
            tr.trim(text)
        assert str(ei.value) == res["error"]
        return
    out, stats, consumed = tr.trim(text)
    assert consumed == len(text)
    fastq_cases.check(case, out.tobytes(), stats, adapters)


def test_streaming_calls_reassemble():
    case = [c for c in CASES if c["label"] == "panel_times3"][0]
    adapters = fastq_cases.adapters_of(case)
    text = case["text"].encode("latin-1")
    tr = fastq.FastqTrimmer(adapters, times=case["times"], chunk_bytes=8192)
    outs, stats, pos = [], tr.new_stats(), 0
    for cut in (10_000, 33_333, 90_001, len(text)):
        out, stats, consumed = tr.trim(text[pos:cut], final=(cut == len(text)), stats=stats)
        outs.append(out.tobytes())
        pos += consumed
    assert pos == len(text)
    fastq_cases.check(case, b"".join(outs), stats, adapters)


def test_one_million_reads_against_match_records():
    n, L = 1_000_000, 150
    reads = synth.synth_se(n, L, seed=synth.seed_for(2, 7), device="cpu").numpy()
    text = synth.fastq_text(reads)
    ad = Adapter(synth.TRUSEQ_R1, BACK, max_error_rate=0.1, min_overlap=3)
    tr = fastq.FastqTrimmer([ad], times=1, max_len=L, chunk_bytes=32 << 20)
    out, stats, consumed = tr.trim(text)
    assert consumed == text.size and stats.records == n
    rec = ad.match_to_batch((reads.reshape(-1), np.arange(n + 1, dtype=np.int64) * L))
    hit = rec["status"] == _abi.ATR_ST_MATCH
    keep = np.where(hit, rec["rstart"].astype(np.int64), L)
    assert stats.with_adapters == int(hit.sum()) and stats.bp_in == n * L and stats.bp_out == int(keep.sum())
    # expected text: header(12) + keep + 3 + keep + 1 bytes per record
    H = 12
    rl = H + 2 * keep + 4
    assert out.size == int(rl.sum())
    starts = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rl, out=starts[1:])
    rec_in = text.reshape(n, -1)
    for i in list(range(0, n, 9973)) + [n - 1]:
        k = int(keep[i])
        exp = bytes(rec_in[i, :H]) + bytes(reads[i, :k]) + b"\n+\n" + b"I" * k + b"\n"
        assert out[starts[i]:starts[i + 1]].tobytes() == exp, i
    # removed-length histogram == what the records say
    removed = (L - rec["rstart"].astype(np.int64))[hit]
    hist = np.bincount(removed, minlength=L + 1)
    assert np.array_equal(stats.errors_back[0].sum(axis=1), hist)


# ---- paired-end ("--aligner insert") ----------------------------------------------------------------------------
PE_CASES = fastq_cases.pe_cases()


@pytest.mark.parametrize("chunk", [0, 4096, 30000])
@pytest.mark.parametrize("case", PE_CASES, ids=[c["label"] for c in PE_CASES])
def test_pe_against_reference_cli(case, chunk):
    a1, a2, ia = fastq_cases.pe_objects(case)
    t1, t2 = case["text1"].encode("latin-1"), case["text2"].encode("latin-1")
    res = case["result"]
    tr = fastq.FastqPairTrimmer(a1, a2, ia, chunk_bytes=chunk, times=case.get("times", 1), mismatch_action=case.get("mismatch_action"),
                                **fastq_cases.merge_kwargs(case), **case.get("read_ops", {}))
    if "error" in res:
        with pytest.raises(fastq.FormatError) as ei:
            tr.trim(t1, t2)
        assert str(ei.value) == res["error"]
        return
    outs, stats, consumed = tr.trim(t1, t2)
    assert consumed == (len(t1), len(t2))
    fastq_cases.pe_check(case, tuple(o.tobytes() for o in outs), stats)


def test_pe_merge_streaming_and_both_kernels(monkeypatch):
    """--merge-overlapping through streaming calls with unequal cuts (merged text and counters add up), and with the
    thread-per-pair merge kernel forced instead of the warp wavefront"""
    case = [c for c in PE_CASES if c["label"] == "merge_correct_liberal"][0]
    a1, a2, ia = fastq_cases.pe_objects(case)
    t1, t2 = case["text1"].encode("latin-1"), case["text2"].encode("latin-1")
    kw = dict(times=case.get("times", 1), mismatch_action=case.get("mismatch_action"), **fastq_cases.merge_kwargs(case), **case.get("read_ops", {}))
    tr = fastq.FastqPairTrimmer(a1, a2, ia, chunk_bytes=8192, **kw)
    parts, stats, p1, p2 = [[], [], []], tr.new_stats(), 0, 0
    for c1, c2 in [(20_000, 9_000), (50_000, 80_000), (len(t1), len(t2))]:
        outs, stats, consumed = tr.trim(t1[p1:c1], t2[p2:c2], final=(c1, c2) == (len(t1), len(t2)), stats=stats)
        for k in range(3):
            parts[k].append(outs[k].tobytes())
        p1 += consumed[0]; p2 += consumed[1]
    assert (p1, p2) == (len(t1), len(t2))
    fastq_cases.pe_check(case, tuple(b"".join(p) for p in parts), stats)
    monkeypatch.setenv("ATR_MERGE_KERNEL", "thread")
    outs, stats, consumed = fastq.FastqPairTrimmer(a1, a2, ia, **kw).trim(t1, t2)
    fastq_cases.pe_check(case, tuple(o.tobytes() for o in outs), stats)


def test_pe_streaming_calls_reassemble():
    case = [c for c in PE_CASES if c["label"] == "ragged_lower"][0]
    a1, a2, ia = fastq_cases.pe_objects(case)
    t1, t2 = case["text1"].encode("latin-1"), case["text2"].encode("latin-1")
    tr = fastq.FastqPairTrimmer(a1, a2, ia, chunk_bytes=8192)
    o1, o2, stats, p1, p2 = [], [], tr.new_stats(), 0, 0
    cuts = [(20_000, 9_000), (50_000, 80_000), (len(t1), len(t2))]
    for c1, c2 in cuts:
        final = (c1, c2) == (len(t1), len(t2))
        outs, stats, consumed = tr.trim(t1[p1:c1], t2[p2:c2], final=final, stats=stats)
        o1.append(outs[0].tobytes()); o2.append(outs[1].tobytes())
        p1 += consumed[0]; p2 += consumed[1]
    assert (p1, p2) == (len(t1), len(t2))
    fastq_cases.pe_check(case, (b"".join(o1), b"".join(o2)), stats)


def test_pe_200k_pairs_against_match_records():
    """the fused path against the separately verified batch calls (InsertAdapterCutter.match_batch) + the reference's
    decision rules restated with numpy"""
    from atropos_b200.modifiers import InsertAdapterCutter
    n, L = 200_000, 150
    r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(3, 5), device="cpu")
    r1, r2 = r1.numpy(), r2.numpy()
    case = {"error_rate": 0.1}
    a1, a2, ia = fastq_cases.pe_objects(case)
    tr = fastq.FastqPairTrimmer(a1, a2, ia, max_len=L, chunk_bytes=16 << 20)
    outs, stats, consumed = tr.trim(synth.fastq_text(r1), synth.fastq_text(r2))
    assert stats.records == n
    offs = np.arange(n + 1, dtype=np.int64) * L
    cutter = InsertAdapterCutter(a1, a2, ia)
    ins, fb1, fb2, need = cutter.match_batch((r1.reshape(-1), offs), (r2.reshape(-1), offs))
    hit = ins["insert"]["status"] == _abi.ATR_ST_MATCH
    m1 = np.where(hit, ins["match1"]["status"], fb1["status"]) == _abi.ATR_ST_MATCH
    m2 = np.where(hit, ins["match2"]["status"], fb2["status"]) == _abi.ATR_ST_MATCH
    s1 = np.where(hit, ins["match1"]["rstart"], fb1["rstart"]).astype(np.int64)
    s2 = np.where(hit, ins["match2"]["rstart"], fb2["rstart"]).astype(np.int64)
    only1, only2 = m1 & ~m2, m2 & ~m1                       # symmetric duplication (equal read lengths here)
    p1 = m1 | (only2 & (s2 <= L))
    p2 = m2 | (only1 & (s1 <= L))
    s1 = np.where(only2, s2, s1)
    s2 = np.where(only1, s1, s2)
    k1 = np.where(p1 & (s1 < L), s1, L)
    k2 = np.where(p2 & (s2 < L), s2, L)
    assert stats.insert_matches == int(hit.sum())
    assert stats.with_adapters == [int(p1.sum()), int(p2.sum())]
    assert stats.bp_out == [int(k1.sum()), int(k2.sum())]
    for out, k in ((outs[0], k1), (outs[1], k2)):
        assert out.size == int((12 + 2 * k + 4).sum())
        seq_lens = np.array([len(x) for x in out.tobytes().split(b"\n")[1::4]], dtype=np.int64)
        assert np.array_equal(seq_lens, k)


def test_file_streaming_helpers(tmp_path):
    """fastq.trim_file / trim_file_pair: blocks far smaller than the files, partial records carried over"""
    case = [c for c in CASES if c["label"] == "ops_quality_trimn_minlen"][0]
    adapters = fastq_cases.adapters_of(case)
    (tmp_path / "in.fq").write_bytes(case["text"].encode("latin-1"))
    tr = fastq.FastqTrimmer(adapters, times=case["times"], **case["read_ops"])
    stats = fastq.trim_file(tr, str(tmp_path / "in.fq"), str(tmp_path / "out.fq"), block_bytes=7001)
    fastq_cases.check(case, (tmp_path / "out.fq").read_bytes(), stats, adapters)
    pcase = [c for c in PE_CASES if c["label"] == "ops_trimn_minlen"][0]
    a1, a2, ia = fastq_cases.pe_objects(pcase)
    (tmp_path / "in1.fq").write_bytes(pcase["text1"].encode("latin-1"))
    (tmp_path / "in2.fq").write_bytes(pcase["text2"].encode("latin-1"))
    ptr = fastq.FastqPairTrimmer(a1, a2, ia, **pcase["read_ops"])
    pstats = fastq.trim_file_pair(ptr, str(tmp_path / "in1.fq"), str(tmp_path / "in2.fq"), str(tmp_path / "o1.fq"),
                                  str(tmp_path / "o2.fq"), block_bytes=9973)
    fastq_cases.pe_check(pcase, ((tmp_path / "o1.fq").read_bytes(), (tmp_path / "o2.fq").read_bytes()), pstats)


def test_compressed_files(tmp_path):
    """trim_file / trim_file_pair on .gz (two members), .bz2 and .xz inputs and a .gz output == the plain-text run"""
    import bz2
    import gzip
    import lzma
    from atropos_b200 import fastq
    case = [c for c in CASES if c["label"] == "ragged_lower_n"][0]
    text = case["text"].encode("latin-1")
    tr = fastq.FastqTrimmer(fastq_cases.adapters_of(case), times=case["times"])
    want = case["result"]["out"].encode("latin-1")
    cut = text.index(b"\n@", len(text) // 2) + 1
    (tmp_path / "a.fq.gz").write_bytes(gzip.compress(text[:cut]) + gzip.compress(text[cut:]))
    (tmp_path / "a.fq.bz2").write_bytes(bz2.compress(text))
    (tmp_path / "a.fq.xz").write_bytes(lzma.compress(text))
    for name in ("a.fq.gz", "a.fq.bz2", "a.fq.xz"):
        st = fastq.trim_file(tr, str(tmp_path / name), str(tmp_path / "out.fq.gz"), block_bytes=20000)
        assert gzip.open(tmp_path / "out.fq.gz").read() == want
        assert st.records == case["result"]["records"] and st.with_adapters == case["result"]["with_adapters"]
