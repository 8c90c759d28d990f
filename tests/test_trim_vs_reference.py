""""Next" row f-1/f-3: the batch consumer of the match records (atropos_b200/trim.py) against the reference's own
AdapterCutter.__call__ + Adapter.trimmed bookkeeping, read by read (needs /root/reference; the match records are
produced by the CPU oracle here -- the GPU produces bit-identical ones, see test_gpu_parity.py)."""
import numpy as np
import pytest

import fuzzgen
from atropos_b200 import _abi, trim
from oracle import oracle

pytestmark = []

SPECS = [("AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC", oracle.BACK), ("TGGAATTCTCGGGTGCCAAGG", oracle.BACK),
         ("GTTCAGAGTTCTACAGTCCGACGATC", oracle.FRONT), ("AATGATACGGCGACCACCGA", oracle.ANYWHERE)]


def _oracle_rounds(reads, adapters, times):
    """What modifiers.AdapterCutter.match_rounds_batch returns, computed with the CPU oracle."""
    n = len(reads)
    cur = list(reads)
    active = [len(r) > 0 for r in reads]
    rounds = []
    for _ in range(times):
        rec = np.zeros(n, dtype=_abi.MATCH_DTYPE)
        rec["adapter"] = -1
        anyhit = False
        for i in range(n):
            if not active[i]:
                continue
            bm = oracle.best_match(adapters, cur[i])
            if bm is None:
                active[i] = False
                continue
            anyhit = True
            a, m = bm
            rec[i] = (m[0], m[1], m[2], m[3], m[4], m[5], a, _abi.ATR_ST_MATCH)
            cur[i] = cur[i][m[3]:] if m[6] else cur[i][:m[2]]
        rounds.append(rec)
        if not anyhit:
            break
    return rounds


def test_trim_and_stats_match_reference(reference):
    from atropos.adapters import Adapter
    from atropos.commands.trim.modifiers import AdapterCutter
    from atropos.io.seqio import Sequence
    rng = np.random.default_rng(2024)
    reads = []
    for _ in range(4000):
        s, w = SPECS[int(rng.integers(0, len(SPECS)))]
        body = fuzzgen.read_with_adapter(rng, s, int(rng.integers(0, 160)), n_rate=0.01)
        if w != oracle.BACK and rng.random() < 0.6:
            body = (fuzzgen.mutate(rng, s, 0.03, 0.01, 0.01) + body)[:150]
        if rng.random() < 0.1:
            body = body.lower()
        reads.append(body)
    times = 2
    ref_adapters = [Adapter(s, w) for s, w in SPECS]
    cutter = AdapterCutter(ref_adapters, times=times, action='trim')
    expected = [cutter(Sequence(name="r%d" % i, sequence=r)).sequence for i, r in enumerate(reads)]

    mine = [oracle.OracleAdapter(s, w) for s, w in SPECS]
    rounds = _oracle_rounds(reads, mine, times)
    blob = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=offs[1:])
    lo, hi, stats, with_adapters = trim.apply_rounds(blob, offs, rounds, [a.front_flag for a in mine])
    out_ascii, out_offs = trim.trimmed_batch(blob, offs, lo, hi)
    got = [bytes(out_ascii[out_offs[i]:out_offs[i + 1]]).decode() for i in range(len(reads))]
    assert got == expected
    assert with_adapters == cutter.with_adapters
    for ra, st in zip(ref_adapters, stats):
        assert dict(ra.lengths_front) == st.lengths_front
        assert dict(ra.lengths_back) == st.lengths_back
        assert {k: dict(v) for k, v in ra.errors_front.items()} == st.errors_front
        assert {k: dict(v) for k, v in ra.errors_back.items()} == st.errors_back
        assert dict(ra.adjacent_bases) == st.adjacent_bases
    assert sum(sum(s.lengths_back.values()) for s in stats) > 500
    assert sum(sum(s.lengths_front.values()) for s in stats) > 200


def test_stats_merge_is_shard_invariant():
    rng = np.random.default_rng(7)
    reads = [fuzzgen.read_with_adapter(rng, SPECS[0][0], 100) for _ in range(600)]
    mine = [oracle.OracleAdapter(*SPECS[0])]
    def run(rs):
        blob = np.frombuffer("".join(rs).encode(), dtype=np.uint8)
        offs = np.zeros(len(rs) + 1, dtype=np.int64)
        np.cumsum([len(r) for r in rs], out=offs[1:])
        return trim.apply_rounds(blob, offs, _oracle_rounds(rs, mine, 1), [False])[2][0]
    whole = run(reads)
    merged = run(reads[:250]).merge(run(reads[250:]))
    assert whole.lengths_back == merged.lengths_back and whole.errors_back == merged.errors_back
    assert whole.adjacent_bases == merged.adjacent_bases
