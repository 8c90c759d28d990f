#!/usr/bin/env python
"""Generate tests/golden/merge_overlap.json.gz from the REAL reference (run in the build container only).

    python tests/golden/make_merge_golden.py

Runs the unmodified reference MergeOverlapping (atropos/commands/trim/modifiers.py:864-931) on the seeded pairs of
tests/fuzzgen.py: merge_cases and stores inputs + outputs: the reads after the call, whether read 2 was dropped, the
correction counters, or the exception type. The alignment the modifier computed inside (Aligner(rc(read2), rate,
flags).locate(read1)) is stored next to it so that the GPU records can be compared field by field.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fuzzgen  # noqa: E402
from make_golden import dump  # noqa: E402
from oracle import build_ref, ref_loader  # noqa: E402


def run_reference(cases):
    from atropos.align import SEMIGLOBAL, START_WITHIN_SEQ1, STOP_WITHIN_SEQ2
    from atropos.align._align import Aligner
    from atropos.commands.trim.modifiers import MergeOverlapping
    from atropos.io.seqio import Sequence
    from atropos.util import reverse_complement
    out = []
    for c in cases:
        mod = MergeOverlapping(min_overlap=c["min_overlap"], error_rate=c["error_rate"], mismatch_action=c["mismatch_action"])
        r1 = Sequence("r", c["seq1"], c["qual1"], insert_overlap=c["insert_matched"])
        r2 = Sequence("r", c["seq2"], c["qual2"], insert_overlap=c["insert_matched"])
        res = {}
        try:
            flags = (START_WITHIN_SEQ1 | STOP_WITHIN_SEQ2) if c["insert_matched"] else SEMIGLOBAL
            res["alignment"] = None
            mo = mod.min_overlap
            if mo <= 1:
                mo = max(2, round(mo * min(len(c["seq1"]), len(c["seq2"]))))
            res["min_overlap"] = mo
            if len(c["seq1"]) >= mo and len(c["seq2"]) >= mo:
                al = Aligner(reverse_complement(c["seq2"]), c["error_rate"], flags).locate(c["seq1"])
                res["alignment"] = None if al is None else list(al)
            a, b = mod(r1, r2)
            res.update(seq1=a.sequence, qual1=a.qualities, merged=bool(a.merged), corrected1=a.corrected,
                       read2=None if b is None else [b.sequence, b.qualities, b.corrected],
                       summary=mod.summarize(), counters=[mod.corrected_pairs, list(mod.corrected_bp)])
        except Exception as e:                              # KeyError (complement), AtroposError (invalid alignment), ...
            res["raises"] = type(e).__name__
        d = dict(c)
        d["result"] = res
        out.append(d)
    return out


def main():
    build_ref.build()
    ref_loader.load_package()
    cases = run_reference(fuzzgen.merge_cases(9101, 2500))
    print("merged:", sum(1 for c in cases if c["result"].get("merged")), "raises:",
          sum(1 for c in cases if "raises" in c["result"]), "of", len(cases))
    dump("merge_overlap.json", cases)


if __name__ == "__main__":
    main()
