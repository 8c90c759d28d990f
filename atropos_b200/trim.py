"""Batch consumption of the match records: trimming windows and the per-adapter statistics of the reference,
without creating one Python object per read ("next" row f-1/f-3 of SURVEY.md section 8: the batch-staged side
of `AdapterCutter.__call__`, atropos/commands/trim/modifiers.py:124-187, and of `Adapter._trimmed_front` /
`_trimmed_back`, atropos/adapters/__init__.py:413-436).

Input: the list of per-round MATCH_DTYPE arrays that `modifiers.AdapterCutter.match_rounds_batch` returns (one
record per read and round; coordinates relative to the window the previous rounds left). Output: the window
`[lo, hi)` of every read that survives the trim, and for every adapter the same dictionaries the reference fills
one read at a time:

    lengths_front[rstop]                      errors_front[rstop][errors]
    lengths_back[len(read) - rstart]          errors_back[len(read) - rstart][errors]
    adjacent_bases[base before the adapter]   (only A, C, G, T are counted by letter, anything else as '')

All of it is numpy on the host: the arithmetic is a handful of integer ops per matched read.
"""
import numpy as np

from . import _abi


class AdapterStats(object):
    """The statistics `Adapter.summarize()` reports (adapters/__init__.py:474-505), as plain dicts."""

    def __init__(self):
        self.lengths_front, self.lengths_back = {}, {}
        self.errors_front, self.errors_back = {}, {}
        self.adjacent_bases = {'A': 0, 'C': 0, 'G': 0, 'T': 0, '': 0}

    def _add(self, lengths, errors, keys, errs):
        if len(keys) == 0:
            return
        pairs, counts = np.unique(np.stack([keys, errs], axis=1), axis=0, return_counts=True)
        for (k, e), c in zip(pairs.tolist(), counts.tolist()):
            lengths[k] = lengths.get(k, 0) + c
            errors.setdefault(k, {})
            errors[k][e] = errors[k].get(e, 0) + c

    def merge(self, other):
        """Sum two shards' statistics (what Summary.merge does for the reference's workers, multicore.py:389)."""
        for mine, theirs in ((self.lengths_front, other.lengths_front), (self.lengths_back, other.lengths_back),
                             (self.adjacent_bases, other.adjacent_bases)):
            for k, v in theirs.items():
                mine[k] = mine.get(k, 0) + v
        for mine, theirs in ((self.errors_front, other.errors_front), (self.errors_back, other.errors_back)):
            for k, d in theirs.items():
                tgt = mine.setdefault(k, {})
                for e, v in d.items():
                    tgt[e] = tgt.get(e, 0) + v
        return self


def apply_rounds(ascii, offsets, rounds, front_flags):
    """Consume the rounds of `AdapterCutter.match_rounds_batch`.

    front_flags: per adapter of the cutter, True (FRONT/PREFIX), False (BACK/SUFFIX) or None (ANYWHERE: front iff
    rstart == 0, align/__init__.py:108-114).
    Returns (lo, hi, stats, with_adapters): int64 windows per read, a list of AdapterStats (one per adapter) and the
    number of reads with at least one match (AdapterCutter.with_adapters)."""
    ascii = np.asarray(ascii, dtype=np.uint8)
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    lo = np.zeros(n, dtype=np.int64)
    hi = np.diff(offsets).astype(np.int64)
    stats = [AdapterStats() for _ in front_flags]
    ff = np.array([-1 if f is None else int(bool(f)) for f in front_flags], dtype=np.int64)
    any_hit = np.zeros(n, dtype=bool)
    letters = {65: 'A', 67: 'C', 71: 'G', 84: 'T'}
    for rec in rounds:
        hit = rec["status"] == _abi.ATR_ST_MATCH
        if not hit.any():
            break
        any_hit |= hit
        idx = np.nonzero(hit)[0]
        ad = rec["adapter"][idx].astype(np.int64)
        rstart = rec["rstart"][idx].astype(np.int64)
        rstop = rec["rstop"][idx].astype(np.int64)
        errs = rec["errors"][idx].astype(np.int64)
        f = ff[ad]
        front = np.where(f < 0, rstart == 0, f == 1)
        wl = hi[idx] - lo[idx]                            # len(match.read): the read as the previous rounds left it
        for a in np.unique(ad).tolist():
            sel_f = (ad == a) & front
            sel_b = (ad == a) & ~front
            st = stats[a]
            st._add(st.lengths_front, st.errors_front, rstop[sel_f], errs[sel_f])
            st._add(st.lengths_back, st.errors_back, (wl - rstart)[sel_b], errs[sel_b])
            if sel_b.any():
                rs = rstart[sel_b]
                pos = offsets[idx[sel_b]] + lo[idx[sel_b]] + rs - 1
                base = np.where(rs >= 1, ascii[np.clip(pos, 0, max(len(ascii) - 1, 0))] if len(ascii) else 0, 0)
                vals, cnts = np.unique(base, return_counts=True)
                for v, c in zip(vals.tolist(), cnts.tolist()):
                    st.adjacent_bases[letters.get(v, '')] += c
        new_lo = np.where(front, lo[idx] + rstop, lo[idx])
        new_hi = np.where(front, hi[idx], lo[idx] + rstart)
        lo[idx], hi[idx] = new_lo, new_hi
    return lo, hi, stats, int(any_hit.sum())


def trimmed_batch(ascii, offsets, lo, hi):
    """Gather the surviving windows into a new contiguous (ascii, offsets) batch."""
    ascii = np.asarray(ascii, dtype=np.uint8)
    offsets = np.asarray(offsets, dtype=np.int64)
    lens = (hi - lo).astype(np.int64)
    new_off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=new_off[1:])
    total = int(new_off[-1])
    if total == 0:
        return np.zeros(0, dtype=np.uint8), new_off
    starts = offsets[:-1] + lo
    src = np.repeat(starts - new_off[:-1], lens) + np.arange(total, dtype=np.int64)
    return ascii[src], new_off
