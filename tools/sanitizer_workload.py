#!/usr/bin/env python
"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitizer_workload.py

K1 funnel both ways (q-gram first stage and Shift-And), the Myers filter path (FRONT), anchored and general kernels, a panel
with rounds, K2 packed + byte kernels, the merge kernels, the packers, the single-end and the paired-end FASTQ paths (with MergeOverlapping) -- each checked against the
oracle or the golden records, so a sanitizer run is also a parity run. Prints 'sanitizer workload ok'."""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from atropos_b200 import _abi, engine, synth
    from atropos_b200.adapters import Adapter, BACK, FRONT, PREFIX, ANYWHERE
    from atropos_b200.align import InsertAligner
    from atropos_b200.modifiers import AdapterCutter
    from oracle import oracle as orc
    import fuzzgen
    rng = np.random.default_rng(5)
    n, L = 3000, 150
    reads = synth.synth_se(n, L, seed=11, device="cpu").numpy()
    offs = np.arange(n + 1, dtype=np.int64) * L
    flat = reads.reshape(-1)
    for seq, where, rate in ((synth.TRUSEQ_R1, BACK, 0.1), ("TGGAATTCTCGGGTGCCAAGG", BACK, 0.1), ("AGATCGGAAGAGC", BACK, 0.1),
                             (synth.TRUSEQ_R2, BACK, 0.1), ("AATGATACGGCGACCACCGA", FRONT, 0.1),
                             ("GTTCAGAGTTCTACAGTCCGACGATC", PREFIX, 0.1), (synth.TRUSEQ_R1, ANYWHERE, 0.2)):
        got = Adapter(seq, where, max_error_rate=rate).match_to_batch((flat, offs))
        oa = orc.OracleAdapter(seq, where, rate, 3)
        for i in range(0, n, 7):
            e = oa.match_to(bytes(reads[i]).decode())
            g = got[i]
            assert (e is None) == (int(g["status"]) == _abi.ATR_ST_NONE), (seq, i)
            if e is not None:
                assert tuple(int(g[k]) for k in ("astart", "astop", "rstart", "rstop", "matches", "errors")) == tuple(e[:6]), (seq, i)
    # ragged reads with lower case / other bytes (general kernel), windows, a panel with 2 rounds
    ragged = [fuzzgen.read_with_adapter(rng, synth.TRUSEQ_R1, int(rng.integers(0, 230)), n_rate=0.02) for _ in range(1500)]
    ragged = [r.lower() if rng.random() < 0.05 else r for r in ragged]
    cutter = AdapterCutter([Adapter(synth.TRUSEQ_R1, BACK), Adapter("AATGATACGGCGACCACCGA", FRONT), Adapter("AGATCGGAAGAGC", BACK)], times=2)
    rounds = cutter.match_rounds_batch(ragged)
    assert len(rounds) >= 1
    codes, woff, lens = engine.pack_reads_host(*engine.encode_reads(ragged), fold_case=True)
    a, o = engine.encode_reads(ragged)
    assert np.array_equal(cutter._adapterset().locate_host_packed(codes, woff, lens, ascii=a, offsets=o, fold_case=True),
                          cutter.best_match_batch((a, o)))
    # K2
    p = 1500
    r1, r2 = synth.synth_pe(p, L, seed=12, device="cpu")
    r1, r2 = r1.numpy(), r2.numpy()
    po = np.arange(p + 1, dtype=np.int64) * L
    for rate in (0.1, 0.2):
        ia = InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, max_insert_mismatch_frac=rate, max_adapter_mismatch_frac=rate)
        res = ia.match_insert_batch((r1.reshape(-1), po), (r2.reshape(-1), po))
        oia = orc.OracleInsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, max_insert_mismatch_frac=rate, max_adapter_mismatch_frac=rate)
        for i in range(0, p, 5):
            e = oia.match_insert(bytes(r1[i]).decode(), bytes(r2[i]).decode())
            g = InsertAligner.result_from_record(res[i])
            assert (e is None) == (g is None) and (e is None or g[0] == e[0]), i
    low = [s.lower() for s in (bytes(x).decode() for x in r1[:200])]
    ia.match_insert_batch(low, [bytes(x).decode() for x in r2[:200]])
    # merge kernels on the golden pairs
    with gzip.open(os.path.join(ROOT, "tests", "golden", "merge_overlap.json.gz"), "rb") as fh:
        cases = [c for c in json.loads(fh.read().decode("ascii")) if c["min_overlap"] == 0.9 and c["error_rate"] == 0.2][:400]
    a1, o1 = engine.encode_reads([c["seq1"].encode("latin-1") for c in cases])
    a2, o2 = engine.encode_reads([c["seq2"].encode("latin-1") for c in cases])
    im = np.array([c["insert_matched"] for c in cases], dtype=np.uint8)
    recs = engine.default_context(0).merge_overlap_host(a1, o1, a2, o2, 0.9, 0.2, insert_matched=im)
    for c, r in zip(cases, recs):
        al = c["result"].get("alignment")
        if al is not None and int(r["status"]) != _abi.ATR_ST_KEYERROR:
            assert [int(r[k]) for k in ("r2_start", "r2_stop", "r1_start", "r1_stop", "matches", "errors")] == al
    # FASTQ paths against the reference CLI's goldens
    from atropos_b200 import fastq
    import fastq_cases
    with gzip.open(os.path.join(ROOT, "tests", "golden", "fastq_trim.json.gz"), "rb") as fh:
        fcases = json.loads(fh.read().decode("ascii"))
    for case in fcases:
        if "error" in case["result"] or case["label"] not in ("se150_truseq", "panel_times3", "no_final_newline_untrimmed", "ops_panel_times2_all"):
            continue
        tr = fastq.FastqTrimmer(fastq_cases.adapters_of(case), times=case["times"], chunk_bytes=20000, **case.get("read_ops", {}))
        out, st, _ = tr.trim(case["text"].encode("latin-1"))
        assert bytes(out) == case["result"]["out"].encode("latin-1"), case["label"]
    # paired-end FASTQ paths: insert mode with correction, adapter mode, and MergeOverlapping with a merged output
    for case in fastq_cases.pe_cases():
        if case["label"] not in ("ragged_lower", "correct_liberal", "adapter_mode_panel_times2_ops", "merge_insert_ragged_ops",
                                 "merge_correct_liberal", "merge_adapter_mode_correct", "merge_discarded"):
            continue
        a1, a2, ia = fastq_cases.pe_objects(case)
        tr = fastq.FastqPairTrimmer(a1, a2, ia, chunk_bytes=30000, times=case.get("times", 1), mismatch_action=case.get("mismatch_action"),
                                    **fastq_cases.merge_kwargs(case), **case.get("read_ops", {}))
        outs, st, _ = tr.trim(case["text1"].encode("latin-1"), case["text2"].encode("latin-1"))
        fastq_cases.pe_check(case, tuple(o.tobytes() for o in outs), st)
    print("sanitizer workload ok")


if __name__ == "__main__":
    main()
