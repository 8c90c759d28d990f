#!/usr/bin/env python
"""Generate tests/golden/*.json.gz from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the unmodified reference package from /root/reference with its Cython modules compiled
into oracle/_ref/ (oracle/build_ref.py), runs it on seeded inputs from tests/fuzzgen.py and
stores inputs + outputs.  The GPU box has no /root/reference; there the committed vectors are
what pins both the CPU oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py).
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fuzzgen  # noqa: E402
from oracle import build_ref, ref_loader  # noqa: E402

TRUSEQ1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"
TRUSEQ2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT"


def dump(name, obj):
    path = os.path.join(HERE, name + ".gz")
    with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as fh:
        fh.write(json.dumps(obj, separators=(",", ":")).encode("ascii"))
    print(name, os.path.getsize(path), "bytes")


def mt(match):
    return None if match is None else [match.astart, match.astop, match.rstart, match.rstop, match.matches,
                                       match.errors, bool(match.front)]


def main():
    build_ref.build()
    ref_loader.load_package()
    from atropos.align import InsertAligner
    from atropos.align._align import Aligner, MultiAligner
    from atropos.adapters import Adapter
    from atropos.io.seqio import Sequence
    from atropos.util import RandomMatchProbability

    # --- Aligner.locate -------------------------------------------------------------------
    cases = []
    for c in fuzzgen.locate_cases(7001, 2500):
        c["wildcard_ref"] = c["wildcard_query"] = False
        cases.append(c)
    cases.extend(fuzzgen.wildcard_locate_cases(7002, 1000))
    # the headline configuration: TruSeq adapter, BACK, 0.1, min_overlap 3, 150 nt reads
    rng = np.random.default_rng(7003)
    for _ in range(1500):
        ad = TRUSEQ1 if rng.random() < 0.7 else "AGATCGGAAGAGC"
        cases.append(dict(reference=ad, query=fuzzgen.read_with_adapter(rng, ad, 150, n_rate=0.002),
                          max_error_rate=0.1, flags=14, wildcard_ref=False, wildcard_query=False, min_overlap=3,
                          indel_cost=1))
    for c in cases:
        r = Aligner(c["reference"], c["max_error_rate"], c["flags"], c["wildcard_ref"], c["wildcard_query"],
                    c["min_overlap"], c["indel_cost"]).locate(c["query"])
        c["expect"] = None if r is None else list(r)
    dump("locate.json", cases)

    # --- MultiAligner.locate (insert flags) -----------------------------------------------
    rng = np.random.default_rng(7004)
    cases = []
    for _ in range(800):
        alpha = rng.choice(["ACGT", "ACGT", "ACGT", "AC", "A"])
        m = int(rng.integers(1, 160))
        ref = fuzzgen.rand_seq(rng, m, alpha)
        ov = int(rng.integers(1, m + 1))
        q = (fuzzgen.mutate(rng, ref[m - ov:], 0.05, 0, 0, alpha) + fuzzgen.rand_seq(rng, m, alpha))[:m] \
            if rng.random() < 0.6 else fuzzgen.rand_seq(rng, m, alpha)
        rate = float(rng.choice([0.0, 0.1, 0.15, 0.2, 0.3]))
        mo = int(rng.choice([1, 3, 10]))
        r = MultiAligner(rate, 9, mo).locate(ref, q)
        cases.append(dict(reference=ref, query=q, max_error_rate=rate, flags=9, min_overlap=mo,
                          expect=None if r is None else [list(t) for t in r]))
    dump("multi_locate.json", cases)

    # --- Adapter.match_to -----------------------------------------------------------------
    rng = np.random.default_rng(7005)
    rmp = RandomMatchProbability()
    cases = []
    wheres = [14, 14, 14, 11, 15, 8, 2]
    for _ in range(500):
        where = wheres[int(rng.integers(0, len(wheres)))]
        wild = rng.random() < 0.25
        seq = fuzzgen.rand_seq(rng, int(rng.integers(3, 50)), "ACGTACGTACGTNRY" if wild else "ACGT")
        kw = dict(max_error_rate=float(rng.choice([0.0, 0.1, 0.12, 0.2])), min_overlap=int(rng.choice([1, 3, 5])),
                  read_wildcards=bool(rng.random() < 0.2), adapter_wildcards=bool(rng.random() < 0.8),
                  indels=bool(rng.random() < 0.8), indel_cost=int(rng.choice([1, 1, 3])))
        max_rmp = [None, 1e-6, 1e-3][int(rng.integers(0, 3))]
        ad = Adapter(seq, where, match_probability=rmp, max_rmp=max_rmp, **kw)
        proj = "".join(ch if ch in "ACGT" else "ACGT"[int(rng.integers(0, 4))] for ch in seq)
        reads, expect = [], []
        for _ in range(4):
            read = fuzzgen.read_with_adapter(rng, proj, int(rng.integers(0, 160)), n_rate=0.01)
            if rng.random() < 0.1:
                read = read.lower()
            reads.append(read)
            expect.append(mt(ad.match_to(Sequence(name="r", sequence=read))))
        cases.append(dict(sequence=seq, where=where, max_rmp=max_rmp, kw=kw, reads=reads, expect=expect))
    dump("match_to.json", cases)

    # --- InsertAligner.match_insert -------------------------------------------------------
    groups = []
    cfgs = [
        dict(max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1),          # BASELINE cfg 3 (-e 0.1)
        dict(max_insert_mismatch_frac=0.2, max_adapter_mismatch_frac=0.2),          # reference default
        dict(max_insert_mismatch_frac=0.15, max_adapter_mismatch_frac=0.15),        # BASELINE cfg 5
        dict(max_insert_mismatch_frac=0.3, max_adapter_mismatch_frac=0.2, min_insert_overlap=10,
             adapter_wildcards=False),
    ]
    for gi, kw in enumerate(cfgs):
        ia = InsertAligner(TRUSEQ1, TRUSEQ2, **kw)
        pairs, expect = [], []
        for r1, r2 in fuzzgen.insert_pairs(7100 + gi, 500, TRUSEQ1, TRUSEQ2, err=[0.01, 0.03, 0.05, 0.08][gi]):
            res = ia.match_insert(r1, r2)
            pairs.append([r1, r2])
            expect.append(None if res is None else [list(res[0]), mt(res[1]), mt(res[2])])
        groups.append(dict(adapter1=TRUSEQ1, adapter2=TRUSEQ2, kw=kw, pairs=pairs, expect=expect))
    dump("match_insert.json", groups)

    # --- RandomMatchProbability -----------------------------------------------------------
    vals = []
    for size in list(range(0, 40)) + [50, 75, 100, 150, 151, 200, 300]:
        for matches in range(0, size + 1, 1 if size < 40 else 5):
            vals.append([matches, size, float(rmp(matches, size)).hex()])
    dump("rmp.json", vals)


if __name__ == "__main__":
    main()
