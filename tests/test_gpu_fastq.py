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
    tr = fastq.FastqTrimmer(adapters, times=case["times"], chunk_bytes=chunk)
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
