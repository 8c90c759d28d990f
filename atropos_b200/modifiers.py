"""Batched front-ends of the two modifiers that call the hot path (reference:
atropos/commands/trim/modifiers.py): `AdapterCutter` (:91-195: best of N adapters, `times` rounds) and
the matching half of `InsertAdapterCutter` (:359-453: insert match first, per-read adapter match as the
fallback). Trimming, masking, error correction and statistics consume the returned records on the host
and stay with the reference (out of scope).
"""
import numpy as np

from . import _abi, engine
from .adapters import Adapter


class AdapterCutter(object):
    """AdapterCutter(adapters, times=1): the alignment part, over a whole batch."""

    def __init__(self, adapters=None, times=1, device=0):
        self.adapters = adapters or []
        self.times = times
        self._device = device
        self._set = None

    def _adapterset(self):
        if self._set is None:
            self._set = engine.AdapterSet(engine.default_context(self._device), [a.descriptor() for a in self.adapters])
        return self._set

    def best_match_batch(self, reads, win=None):
        """_best_match (modifiers.py:107-122) for every read: one MATCH_DTYPE record per read; `adapter` is the
        index into self.adapters of the winner (strictly more matches wins, the first adapter on ties)."""
        ascii, offsets = reads if isinstance(reads, tuple) else engine.encode_reads(reads)
        return self._adapterset().locate_host(ascii, offsets, win=win, fold_case=True)

    def match_rounds_batch(self, reads):
        """__call__'s `for _ in range(self.times)` loop (:143-149) over a batch: returns a list of up to `times`
        MATCH_DTYPE arrays; round t+1 aligns inside the window that round t's trim leaves (front adapters cut
        [0, rstop), back adapters cut [rstart, len)); coordinates of every round are relative to its window,
        like the reference's coordinates are relative to the already-trimmed read."""
        ascii, offsets = reads if isinstance(reads, tuple) else engine.encode_reads(reads)
        n = len(offsets) - 1
        lens = np.diff(offsets).astype(np.int64)
        lo = np.zeros(n, dtype=np.int64)
        hi = lens.copy()
        active = hi > lo                         # `if len(read) == 0: return read` (:136-137)
        front_flags = np.array([-1 if a._front_flag is None else int(a._front_flag) for a in self.adapters])
        rounds = []
        for _ in range(self.times):
            win = np.stack([lo, hi], axis=1).astype(np.uint16)
            res = self.best_match_batch((ascii, offsets), win=win)
            res["status"][~active] = _abi.ATR_ST_NONE
            hit = res["status"] == _abi.ATR_ST_MATCH
            rounds.append(res)
            if not hit.any():
                break
            ff = front_flags[np.clip(res["adapter"], 0, None)]
            is_front = np.where(ff < 0, res["rstart"] == 0, ff == 1)
            # Adapter._trimmed_front keeps read[rstop:], _trimmed_back keeps read[:rstart] (adapters :413-436)
            new_lo = np.where(hit & is_front, lo + res["rstop"], lo)
            new_hi = np.where(hit & ~is_front, lo + res["rstart"], hi)
            lo, hi = new_lo, new_hi
            active = hit
        return rounds


class InsertAdapterCutter(object):
    """The matching half of InsertAdapterCutter.__call__ (modifiers.py:391-415) over a batch of pairs."""

    def __init__(self, adapter1, adapter2, insert_aligner, min_insert_overlap=1):
        assert isinstance(adapter1, Adapter) and isinstance(adapter2, Adapter)
        self.adapter1, self.adapter2 = adapter1, adapter2
        self.aligner = insert_aligner
        self.min_insert_overlap = min_insert_overlap

    def match_batch(self, reads1, reads2):
        """Returns (insert INSERT_DTYPE array, fallback1 MATCH_DTYPE, fallback2 MATCH_DTYPE, used_fallback mask).
        Pairs with a read shorter than min_insert_overlap are skipped (:392-394); pairs without an insert match
        get adapter1.match_to(read1) / adapter2.match_to(read2) (:401-406)."""
        a1, o1 = reads1 if isinstance(reads1, tuple) else engine.encode_reads(reads1)
        a2, o2 = reads2 if isinstance(reads2, tuple) else engine.encode_reads(reads2)
        ins = self.aligner.match_insert_batch((a1, o1), (a2, o2))
        l1, l2 = np.diff(o1), np.diff(o2)
        skipped = (l1 < self.min_insert_overlap) | (l2 < self.min_insert_overlap)
        ins["insert"]["status"][skipped] = _abi.ATR_ST_NONE
        need = (ins["insert"]["status"] == _abi.ATR_ST_NONE) & ~skipped
        # the per-read fallback only runs where it is needed: everything else gets an empty window (no DP work)
        win1 = np.zeros((len(l1), 2), dtype=np.uint16)
        win2 = np.zeros((len(l2), 2), dtype=np.uint16)
        win1[need, 1] = l1[need]
        win2[need, 1] = l2[need]
        fb1 = self.adapter1.match_to_batch((a1, o1), win=win1)
        fb2 = self.adapter2.match_to_batch((a2, o2), win=win2)
        fb1["status"][~need] = _abi.ATR_ST_NONE
        fb2["status"][~need] = _abi.ATR_ST_NONE
        return ins, fb1, fb2, need
