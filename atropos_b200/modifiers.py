"""Batched front-ends of the two modifiers that call the hot path (reference:
atropos/commands/trim/modifiers.py): `AdapterCutter` (:91-195: best of N adapters, `times` rounds) and
the matching half of `InsertAdapterCutter` (:359-453: insert match first, per-read adapter match as the
fallback). Trimming, masking, error correction and statistics consume the returned records on the host
and stay with the reference (out of scope).
"""
import numpy as np

from . import _abi, engine
from .adapters import Adapter
from .util import BASE_COMPLEMENTS, reverse_complement


class AdapterCutter(object):
    """AdapterCutter(adapters, times=1): the alignment part, over a whole batch."""

    def __init__(self, adapters=None, times=1, device=0):
        self.adapters = adapters or []
        self.times = times
        self._device = device
        self._set = None

    def _adapterset(self):
        key = engine.context_key(self._device)
        if self._set is None or self._set_key != key:
            self._set = engine.AdapterSet(engine.default_context(self._device), [a.descriptor() for a in self.adapters])
            self._set_key = key
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


class MergeError(Exception):
    """The reference's AtroposError("Invalid alignment while trying to merge read ...") (modifiers.py:923-927)."""


class MergeOverlapping(object):
    """MergeOverlapping(min_overlap=0.9, error_rate=0.1, mismatch_action=None) (modifiers.py:864-931) over a batch of
    pairs. The per-pair alignment of read 1 against reverse_complement(read 2) and the decision which of the four
    ways the pair is merged run on the GPU (atr_merge_overlap_batch_host); the host only puts the merged strings
    together from the returned records. Reads are objects with `sequence`, `qualities` and, optionally,
    `insert_overlap`, `corrected` and `merged` (the reference's Sequence, io/_seqio.pyx:17-60)."""

    def __init__(self, min_overlap=0.9, error_rate=0.1, mismatch_action=None, device=0):
        self.min_overlap = int(min_overlap) if min_overlap > 1 else min_overlap
        self.error_rate = error_rate
        self.mismatch_action = mismatch_action
        self.r1r2_min_qual_difference = 1
        self.r2r1_min_qual_difference = -1
        self.corrected_pairs = 0
        self.corrected_bp = [0, 0]
        self._device = device

    def align_batch(self, reads1, reads2, insert_matched=None):
        """MERGE_DTYPE records (alignment of read 1 in rc(read 2), effective min_overlap, status, action)."""
        a1, o1 = reads1 if isinstance(reads1, tuple) else engine.encode_reads(reads1)
        a2, o2 = reads2 if isinstance(reads2, tuple) else engine.encode_reads(reads2)
        return engine.default_context(self._device).merge_overlap_host(a1, o1, a2, o2, self.min_overlap, self.error_rate,
                                                                      insert_matched=insert_matched)

    def merge_batch(self, reads1, reads2):
        """[(read1, read2 or None)] like calling the reference modifier on every pair (reads are modified in place)."""
        im = np.fromiter((bool(getattr(a, "insert_overlap", False) and getattr(b, "insert_overlap", False))
                          for a, b in zip(reads1, reads2)), dtype=np.uint8, count=len(reads1))
        recs = self.align_batch([r.sequence for r in reads1], [r.sequence for r in reads2], insert_matched=im)
        return [self.apply_record(a, b, rec, bool(m)) for a, b, rec, m in zip(reads1, reads2, recs, im)]

    def __call__(self, read1, read2):
        return self.merge_batch([read1], [read2])[0]

    def apply_record(self, read1, read2, rec, insert_matched):
        status = int(rec["status"])
        if status == _abi.ATR_ST_NONE:
            return (read1, read2)
        if status == _abi.ATR_ST_KEYERROR:
            reverse_complement(read2.sequence)              # raises the reference's KeyError
        alignment = tuple(int(rec[f]) for f in ("r2_start", "r2_stop", "r1_start", "r1_stop", "matches", "errors"))
        if status == _abi.ATR_ST_INVALID:
            raise MergeError("Invalid alignment while trying to merge read {}: {}".format(
                getattr(read1, "name", ""), ",".join(str(i) for i in alignment)))
        r2_stop, r1_stop, errors = alignment[1], alignment[3], alignment[5]
        read2_rc = reverse_complement(read2.sequence)       # taken before any correction, like the reference (:892)
        if self.mismatch_action and errors > 0 and not insert_matched:
            self.correct_errors(read1, read2, alignment)
        action = int(rec["action"])
        both_quals = read1.qualities and read2.qualities
        if action == _abi.ATR_MERGE_TAKE2:
            read1.sequence = read2_rc
            read1.qualities = "".join(reversed(read2.qualities))
        elif action == _abi.ATR_MERGE_APPEND:
            read1.sequence += read2_rc[r2_stop:]
            if both_quals:
                read1.qualities += "".join(reversed(read2.qualities))[r2_stop:]
        elif action == _abi.ATR_MERGE_PREPEND:
            read1.sequence = read2_rc + read1.sequence[r1_stop:]
            if both_quals:
                read1.qualities = "".join(reversed(read2.qualities)) + read1.qualities[r1_stop:]
        read1.merged = True
        return (read1, None)

    def correct_errors(self, read1, read2, alignment):
        """ErrorCorrectorMixin.correct_errors(read1, read2, insert_match) without truncation (modifiers.py:218-350):
        position t of the aligned stretch of read 1 is paired with position t (from the right) of read 2's stretch,
        whatever indels the alignment has, exactly as the reference pairs them."""
        if getattr(read1, "corrected", 0) > 0 or getattr(read2, "corrected", 0) > 0:
            return
        seq = [list(read1.sequence), list(read2.sequence)]
        has_quals = bool(read1.qualities and read2.qualities)
        if has_quals:
            qual = [list(read1.qualities), list(read2.qualities)]
        elif self.mismatch_action in ("liberal", "conservative"):
            raise ValueError("Cannot perform quality-based error correction on reads lacking quality information")
        len2 = len(seq[1])
        lo1, hi1 = alignment[2], alignment[3]
        lo2, hi2 = len2 - alignment[1], len2 - alignment[0]
        changed = [0, 0]
        undecided = []

        def take(dst, i, j):
            """copy the base of the other read (complemented) and its quality over position i (read 1) / j (read 2)"""
            if dst == 0:
                seq[0][i] = BASE_COMPLEMENTS[seq[1][j]]
                if has_quals:
                    qual[0][i] = qual[1][j]
            else:
                seq[1][j] = BASE_COMPLEMENTS[seq[0][i]]
                if has_quals:
                    qual[1][j] = qual[0][i]
            changed[dst] += 1

        for i, j in zip(range(lo1, hi1), range(hi2 - 1, lo2 - 1, -1)):
            base1, base2 = seq[0][i], BASE_COMPLEMENTS[seq[1][j]]
            if base1 == base2:
                continue
            if self.mismatch_action == "N":
                seq[0][i] = seq[1][j] = "N"
                changed[0] += 1
                changed[1] += 1
            elif base1 == "N":
                take(0, i, j)
            elif base2 == "N":
                take(1, i, j)
            elif has_quals:
                diff = ord(qual[0][i]) - ord(qual[1][j])
                if diff >= self.r1r2_min_qual_difference:
                    take(1, i, j)
                elif diff <= self.r2r1_min_qual_difference:
                    take(0, i, j)
                elif self.mismatch_action == "liberal":
                    undecided.append((i, j))
        if undecided:
            q1 = [ord(c) for c in qual[0][lo1:hi1]]
            q2 = [ord(c) for c in qual[1][lo2:hi2]]
            diff = sum(q1) / len(q1) - sum(q2) / len(q2)
            if diff > 1 or diff < -1:
                for i, j in undecided:
                    take(1 if diff > 1 else 0, i, j)
        if changed[0] or changed[1]:
            self.corrected_pairs += 1
            for k, read in enumerate((read1, read2)):
                if changed[k]:
                    self.corrected_bp[k] += changed[k]
                    read.corrected = changed[k]
                    read.sequence = "".join(seq[k])
                    if has_quals:
                        read.qualities = "".join(qual[k])

    def summarize(self):
        """{}: in the reference ReadPairModifier.summarize comes first in the MRO (modifiers.py:66-69, :864), the
        correction counters stay in `corrected_pairs` / `corrected_bp`."""
        return {}
