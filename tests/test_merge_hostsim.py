"""MergeOverlapping on the CPU build box: the per-pair device function (csrc/merge_core.cuh: merge_pair) compiled for
the host, and the host-side assembly of the merged reads (atropos_b200/modifiers.py), against golden vectors produced
by the reference modifier (atropos/commands/trim/modifiers.py:864-931)."""
import numpy as np

import hostsim
import merge_cases
from atropos_b200 import _abi


def _rec(out):
    rec = np.zeros(1, dtype=_abi.MERGE_DTYPE)[0]
    for f in _abi.MERGE_DTYPE.names:
        rec[f] = getattr(out, f)
    return rec


def test_merge_golden_hostsim():
    cs = merge_cases.cases()
    merged = 0
    for c in cs:
        out = hostsim.merge_overlap(c["seq1"], c["seq2"], c["insert_matched"], c["min_overlap"], c["error_rate"])
        rec = _rec(out)
        merge_cases.check_record(c, rec)
        merge_cases.check_apply(c, rec)
        merged += int(rec["status"]) == _abi.ATR_ST_MATCH
    assert merged > 400


def test_merge_min_overlap_rule():
    # int(min_overlap) if > 1, else a fraction of the shorter read with Python's round (half to even), at least 2
    for mo, l1, l2 in [(0.9, 100, 80), (0.5, 5, 9), (0.5, 7, 7), (0.25, 10, 10), (1.5, 40, 30), (1.0, 33, 35),
                       (12.0, 50, 50), (0.05, 10, 10)]:
        v = int(mo) if mo > 1 else mo
        exp = max(2, round(v * min(l1, l2))) if v <= 1 else v
        out = hostsim.merge_overlap("A" * l1, "C" * l2, False, mo, 0.1)
        assert out.min_overlap == exp, (mo, l1, l2)
