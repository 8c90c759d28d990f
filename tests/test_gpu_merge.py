"""MergeOverlapping through the C ABI (atr_merge_overlap_batch_host -> k_merge_overlap) on the GPU: bit-exact against
golden vectors from the reference modifier (atropos/commands/trim/modifiers.py:864-931), both column placements
(shared memory / global scratch), and a size-independent property at 1 M pairs: two error-free reads of one fragment
merge back into the fragment."""
import os

import numpy as np
import pytest

import merge_cases
from atropos_b200 import _abi, engine
from atropos_b200.modifiers import MergeOverlapping

pytestmark = pytest.mark.gpu


def _groups(cs):
    """cases that share (min_overlap, error_rate) go through one call"""
    groups = {}
    for i, c in enumerate(cs):
        groups.setdefault((c["min_overlap"], c["error_rate"]), []).append(i)
    return groups


def _run(cs, idx, mo, er):
    ctx = engine.default_context(0)
    a1, o1 = engine.encode_reads([cs[i]["seq1"].encode("latin-1") for i in idx])
    a2, o2 = engine.encode_reads([cs[i]["seq2"].encode("latin-1") for i in idx])
    im = np.array([cs[i]["insert_matched"] for i in idx], dtype=np.uint8)
    return ctx.merge_overlap_host(a1, o1, a2, o2, mo, er, insert_matched=im)


@pytest.mark.parametrize("kernel", ["warp", "thread"])
@pytest.mark.parametrize("short_only", [False, True])
def test_merge_golden(short_only, kernel, monkeypatch):
    """kernel: k_merge_warp (one warp per pair, 5 or 10 rows per lane) / k_merge_overlap (one thread per pair, columns
    in shared memory or global scratch) -- the library picks the first whenever its limits hold"""
    monkeypatch.setenv("ATR_MERGE_KERNEL", kernel)
    cs = merge_cases.cases()
    if short_only:                                  # read 2 <= 150 nt: 5 rows per lane / the DP columns in shared memory
        cs = [c for c in cs if len(c["seq2"]) <= 150]
    ctx = engine.default_context(0)
    before = ctx.launch_count()
    merged = 0
    for (mo, er), idx in _groups(cs).items():
        recs = _run(cs, idx, mo, er)
        for i, rec in zip(idx, recs):
            merge_cases.check_record(cs[i], rec)
            merge_cases.check_apply(cs[i], rec)
            merged += int(rec["status"]) == _abi.ATR_ST_MATCH
    assert merged > (300 if short_only else 400)
    assert ctx.launch_count() > before


def test_merge_modifier_batch():
    """the reference-shaped front end: MergeOverlapping.merge_batch on read objects"""
    cs = [c for c in merge_cases.cases() if c["min_overlap"] == 0.9 and c["error_rate"] == 0.2 and c["mismatch_action"] is None
          and "raises" not in c["result"]]
    assert len(cs) > 20
    mod = MergeOverlapping(min_overlap=0.9, error_rate=0.2)
    r1 = [merge_cases.Read(c["seq1"], c["qual1"], c["insert_matched"]) for c in cs]
    r2 = [merge_cases.Read(c["seq2"], c["qual2"], c["insert_matched"]) for c in cs]
    for c, (a, b) in zip(cs, mod.merge_batch(r1, r2)):
        res = c["result"]
        assert (a.sequence, a.qualities, bool(a.merged)) == (res["seq1"], res["qual1"], res["merged"])
        assert (None if b is None else [b.sequence, b.qualities, b.corrected]) == res["read2"]


def test_merge_empty_and_limits():
    ctx = engine.default_context(0)
    z = np.zeros(1, dtype=np.int64)
    assert len(ctx.merge_overlap_host(np.zeros(0, np.uint8), z, np.zeros(0, np.uint8), z, 0.9, 0.1)) == 0
    a, o = engine.encode_reads([b"A" * 4001])
    with pytest.raises(OverflowError):
        ctx.merge_overlap_host(a, o, a, o, 0.9, 0.1)
    with pytest.raises(ValueError):
        ctx.merge_overlap_host(a, o, a, o, 0.0, 0.1)


@pytest.mark.parametrize("L,n", [(150, 1_000_000), (300, 200_000)])
def test_merge_fragments_roundtrip(L, n):
    """read 1 = fragment[:L], read 2 = rc(fragment)[:L] without errors, fragment length F in [L, 2L - 20]: the only
    error-free overlap is the true one (2L - F bases), so every pair must come back as read 1 + rc(read 2)[2L - F:]
    = the fragment, with r1_start = F - L."""
    rng = np.random.default_rng(4242 + L)
    F = rng.integers(L, 2 * L - 20 + 1, size=n)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    frag = np.empty((n, 2 * L), dtype=np.uint8)
    r1 = np.empty((n, L), dtype=np.uint8)
    r2 = np.empty((n, L), dtype=np.uint8)
    for c0 in range(0, n, 250_000):
        c1 = min(n, c0 + 250_000)
        frag[c0:c1] = rng.integers(0, 4, size=(c1 - c0, 2 * L), dtype=np.uint8)
        r1[c0:c1] = acgt[frag[c0:c1, :L]]
        col = (F[c0:c1, None] - 1 - np.arange(L)[None, :]).astype(np.int32)   # read 2 base t = complement of fragment[F - 1 - t]
        r2[c0:c1] = acgt[3 - np.take_along_axis(frag[c0:c1], col, axis=1)]
    offs = engine.fixed_length_offsets(n, L)
    ctx = engine.default_context(0)
    recs = ctx.merge_overlap_host(r1.reshape(-1), offs, r2.reshape(-1), offs, 20, 0.1)
    ov = 2 * L - F
    assert np.all(recs["status"] == _abi.ATR_ST_MATCH)
    assert np.array_equal(recs["r1_start"], F - L) and np.all(recs["r1_stop"] == L)
    assert np.all(recs["r2_start"] == 0) and np.array_equal(recs["r2_stop"], ov)
    assert np.array_equal(recs["matches"], ov) and np.all(recs["errors"] == 0)
    full = F == L                                                       # complete overlap: read 2 lies inside read 1
    assert np.all(recs["action"][full] == _abi.ATR_MERGE_KEEP1) and np.all(recs["action"][~full] == _abi.ATR_MERGE_APPEND)
    # spot-check the assembled strings
    mod = MergeOverlapping(min_overlap=20, error_rate=0.1)
    for i in rng.integers(0, n, size=200):
        a = merge_cases.Read(r1[i].tobytes().decode(), "I" * L, False)
        b = merge_cases.Read(r2[i].tobytes().decode(), "I" * L, False)
        a, b = mod.apply_record(a, b, recs[i], False)
        assert b is None and a.sequence == acgt[frag[i, :F[i]]].tobytes().decode() and len(a.qualities) == F[i]


def test_merge_long_reads_fall_back():
    """reads beyond the warp kernel's 320 nt take the thread-per-pair kernel: same answers as for the same pair inside
    the limits (an exact overlap of 40 bases between two 400 nt reads)"""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    frag = acgt[rng.integers(0, 4, size=760)].tobytes()
    a, o1 = engine.encode_reads([frag[:400]])
    b, o2 = engine.encode_reads([frag[360:].translate(comp)[::-1]])
    rec = engine.default_context(0).merge_overlap_host(a, o1, b, o2, 20, 0.1)[0]
    assert tuple(int(rec[f]) for f in merge_cases.FIELDS) == (0, 40, 360, 400, 40, 0)
    assert (int(rec["status"]), int(rec["action"])) == (_abi.ATR_ST_MATCH, _abi.ATR_MERGE_APPEND)


def test_merge_mixed_lengths_complete_overlap():
    """1 M pairs of mixed lengths (60..150 nt), read 2 = reverse complement of read 1: every pair overlaps completely
    (matches = length, no errors, read 2 inside read 1). Mixed lengths go through the host's ordering by rows per lane."""
    rng = np.random.default_rng(3)
    n = 1_000_000
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    lens = rng.integers(60, 151, size=n)
    offs = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    a1 = acgt[rng.integers(0, 4, size=int(offs[-1]))]
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    idx = np.repeat(offs[:-1] + lens - 1, lens) - (np.arange(int(offs[-1])) - np.repeat(offs[:-1], lens))
    a2 = comp[a1[idx]]
    recs = engine.default_context(0).merge_overlap_host(a1, offs, a2, offs, 0.9, 0.2)
    assert np.all(recs["status"] == _abi.ATR_ST_MATCH) and np.all(recs["action"] == _abi.ATR_MERGE_KEEP1)
    assert np.array_equal(recs["matches"], lens) and np.all(recs["errors"] == 0)
    assert np.all(recs["r1_start"] == 0) and np.array_equal(recs["r1_stop"], lens)
    assert np.all(recs["r2_start"] == 0) and np.array_equal(recs["r2_stop"], lens)


@pytest.mark.parametrize("L,n,rate", [(150, 400_000, 0.2), (100, 300_000, 0.1), (300, 60_000, 0.2)])
def test_merge_warp_equals_thread_kernel(L, n, rate, monkeypatch):
    """the wavefront kernel (every cell, costs clamped) against the thread-per-pair kernel (the reference's banded loop,
    the code the CPU simulator checks against the golden vectors) on pairs with substitutions, indels and Ns:
    identical records, both flag sets"""
    from merge_probe import pairs
    r1, r2, F = pairs(n, L, 900 + L)
    rng = np.random.default_rng(L)
    for arr in (r1, r2):                                   # a few Ns and a deletion-like shift in some reads
        hit = rng.random(arr.shape) < 0.002
        arr[hit] = ord("N")
    shift = rng.random(n) < 0.1
    r2[shift, 40:-1] = r2[shift, 41:]
    offs = engine.fixed_length_offsets(n, L)
    im = (rng.random(n) < 0.3).astype(np.uint8)
    ctx = engine.default_context(0)
    monkeypatch.setenv("ATR_MERGE_KERNEL", "warp")
    a = ctx.merge_overlap_host(r1.reshape(-1), offs, r2.reshape(-1), offs, 0.5, rate, insert_matched=im)
    monkeypatch.setenv("ATR_MERGE_KERNEL", "thread")
    b = ctx.merge_overlap_host(r1.reshape(-1), offs, r2.reshape(-1), offs, 0.5, rate, insert_matched=im)
    assert a.tobytes() == b.tobytes()
    assert (a["status"] == _abi.ATR_ST_MATCH).mean() > 0.1 and (a["errors"] > 0).mean() > 0.1
