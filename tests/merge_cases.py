"""Shared by the CPU (host simulator) and GPU tests of MergeOverlapping: golden cases produced by the reference
modifier (tests/golden/make_merge_golden.py) and the comparison of a record / a merged pair with them."""
import golden_util
from atropos_b200 import _abi
from atropos_b200.modifiers import MergeOverlapping

RAISES = {"AtroposError": "MergeError"}
FIELDS = ("r2_start", "r2_stop", "r1_start", "r1_stop", "matches", "errors")


class Read(object):
    """the attributes of the reference's Sequence that the modifier touches (io/_seqio.pyx:17-60)"""

    def __init__(self, sequence, qualities, insert_overlap):
        self.name = "r"
        self.sequence, self.qualities = sequence, qualities
        self.insert_overlap, self.merged, self.corrected = insert_overlap, False, 0


def cases():
    return golden_util.load("merge_overlap")


def check_record(case, rec):
    """the GPU's (or the simulator's) record against the alignment the reference computed inside the modifier"""
    res = case["result"]
    fields = [int(rec[f]) for f in FIELDS]
    status = int(rec["status"])
    assert int(rec["min_overlap"]) == res["min_overlap"], case
    al = res.get("alignment")
    if status == _abi.ATR_ST_KEYERROR:                       # reverse_complement(read 2) raised before any alignment
        assert res.get("raises") == "KeyError" and al is None, case
        return
    if al is None:
        assert status == _abi.ATR_ST_NONE and fields == [0] * 6, (case, fields, status)
    else:
        assert fields == al, (case, fields, status)
        assert (status == _abi.ATR_ST_NONE) == (al[4] < res["min_overlap"]), (case, status)


def check_apply(case, rec):
    """apply the record like modifiers.MergeOverlapping does and compare the pair with the reference's"""
    res = case["result"]
    mod = MergeOverlapping(min_overlap=case["min_overlap"], error_rate=case["error_rate"], mismatch_action=case["mismatch_action"])
    r1 = Read(case["seq1"], case["qual1"], case["insert_matched"])
    r2 = Read(case["seq2"], case["qual2"], case["insert_matched"])
    try:
        a, b = mod.apply_record(r1, r2, rec, case["insert_matched"])
    except Exception as e:
        assert RAISES.get(res.get("raises"), res.get("raises")) == type(e).__name__, (case, repr(e))
        return
    assert "raises" not in res, case
    assert (a.sequence, a.qualities, bool(a.merged), a.corrected) == (res["seq1"], res["qual1"], res["merged"], res["corrected1"]), case
    assert (None if b is None else [b.sequence, b.qualities, b.corrected]) == res["read2"], case
    assert mod.summarize() == res["summary"], case
    assert [mod.corrected_pairs, list(mod.corrected_bp)] == res["counters"], case
