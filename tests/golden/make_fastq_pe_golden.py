#!/usr/bin/env python
"""Generate tests/golden/fastq_trim_pe.json.gz from the REAL reference command line (build container only).

    python tests/golden/make_fastq_pe_golden.py

Paired-end twin of make_fastq_golden.py: every case is a pair of small FASTQ texts run through the unmodified

    atropos trim --aligner insert -a ADAPTER1 -A ADAPTER2 -pe1 in1.fq -pe2 in2.fq -o out1.fq -p out2.fq

(PairedSequenceReader io/seqio.py:397-453 -> InsertAdapterCutter commands/trim/modifiers.py:359-496 ->
FastqFormat io/seqio.py:686-700). Stored: both output texts and the report's statistics, or the FormatError.
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from atropos_b200 import synth  # noqa: E402
from oracle import build_ref, ref_loader  # noqa: E402

A1, A2 = synth.TRUSEQ_R1, synth.TRUSEQ_R2


def quals(rng, n):
    return "".join(chr(int(c)) for c in rng.integers(33, 74, n))


def fastq(records, eol="\n"):
    return "".join("@" + name + eol + seq + eol + "+" + name2 + eol + q + eol for name, seq, name2, q in records)


def make_pairs(rng, n, L=150, ragged=False, lower=0.0, suffix=False, seed=1):
    r1, r2 = synth.synth_pe(n, L, seed=seed, device="cpu")
    r1, r2 = r1.numpy(), r2.numpy()
    recs1, recs2 = [], []
    for i in range(n):
        s1, s2 = bytes(r1[i]).decode(), bytes(r2[i]).decode()
        if ragged:
            if rng.random() < 0.4:
                s1 = s1[:int(rng.integers(0, L + 1))]
            if rng.random() < 0.4:
                s2 = s2[:int(rng.integers(0, L + 1))]
        if rng.random() < lower:
            s1 = s1.lower()
        if rng.random() < lower:
            s2 = s2.lower()
        base = "pair%d" % i
        c = " 1:N:0:%d" % int(rng.integers(0, 99)) if rng.random() < 0.5 else ""
        n1, n2 = (base + "/1" + c, base + "/2" + c) if suffix else (base + c, base + c.replace(" 1:", " 2:"))
        recs1.append((n1, s1, n1 if rng.random() < 0.1 else "", quals(rng, len(s1))))
        recs2.append((n2, s2, "", quals(rng, len(s2))))
    return recs1, recs2


def adapter_stats(st):
    d = {"sequence": st["sequence"], "where": st["where"]["name"]}
    for key in ("lengths_front", "lengths_back", "adjacent_bases"):
        if key in st:
            d[key] = st[key]
    for key in ("errors_front", "errors_back"):
        if key in st:
            cols = st[key]["columns"]
            d[key] = {ln: {str(c): v for c, v in zip(cols, row) if v} for ln, row in st[key]["rows"].items()}
    return d


def ops_stats(rj):
    t = rj["trim"]
    mods, flt = t["modifiers"], t.get("filters", {})
    bp = lambda name: [int(x or 0) for x in mods[name]["bp_trimmed"]] if name in mods else None
    nf = lambda name: flt[name]["records_filtered"] if name in flt else None
    return {"bp_cut": bp("UnconditionalCutter"), "bp_quality": bp("QualityTrimmer"), "bp_n_ends": bp("NEndTrimmer"),
            "bp_nextseq": bp("NextseqQualityTrimmer"),
            "too_short": nf("too_short"), "too_long": nf("too_long"), "too_many_n": nf("too_many_n"),
            "discarded_trimmed": nf("TrimmedFilter"), "discarded_untrimmed": nf("UntrimmedFilter"),
            "records_written": t["formatters"]["records_written"]}


def run_reference(text1, text2, error_rate, extra=(), adapter_args=None, merge=None):
    from atropos.commands import get_command
    tmp = tempfile.mkdtemp(prefix="fqpegold")
    try:
        p = {k: os.path.join(tmp, k) for k in ("in1.fq", "in2.fq", "out1.fq", "out2.fq", "merged.fq", "rep")}
        for key, text in (("in1.fq", text1), ("in2.fq", text2)):
            with open(p[key], "w", newline="") as fh:
                fh.write(text)
        args = (["--aligner", "insert", "-a", A1, "-A", A2] if adapter_args is None else list(adapter_args)) + ["-pe1", p["in1.fq"], "-pe2", p["in2.fq"], "-o", p["out1.fq"],
                "-p", p["out2.fq"], "--no-default-adapters", "--no-cache-adapters", "--quiet", "--report-file", p["rep"],
                "--report-formats", "json"] + list(extra)
        if error_rate is not None:
            args += ["-e", repr(error_rate)]
        if merge is not None:                          # --merge-overlapping [--merged-output] (trim/__init__.py:546-552, :576-579)
            args += ["--merge-overlapping"]
            if merge.get("output", True):
                args += ["--merged-output", p["merged.fq"]]
            if "min_overlap" in merge:
                args += ["--merge-min-overlap", repr(merge["min_overlap"])]
            if "error_rate" in merge:
                args += ["--merge-error-rate", repr(merge["error_rate"])]
        rc, summary = get_command("trim").execute(args)
        if rc != 0:
            from atropos.io.seqio import FormatError, PairedSequenceReader
            try:
                reader = PairedSequenceReader(p["in1.fq"], p["in2.fq"])
                for _ in reader:
                    pass
            except FormatError as exc:
                return {"error": str(exc)}
            return {"exception": "trim returned %r but the reader raised nothing" % (rc,)}
        corrected = None
        outs = []
        for key in ("out1.fq", "out2.fq"):
            with open(p[key], "r", newline="") as fh:
                outs.append(fh.read())
        rep = p["rep"] + ".json" if os.path.exists(p["rep"] + ".json") else p["rep"]
        with open(rep) as fh:
            rj = json.load(fh)
        if adapter_args is None:
            cutter = rj["trim"]["modifiers"]["InsertAdapterCutter"]
            ads = [adapter_stats(list(d.values())[0]) for d in cutter["adapters"]]
            if "records_corrected" in cutter:
                corrected = {"records_corrected": cutter["records_corrected"], "bp_corrected": cutter["bp_corrected"]}
        else:                                          # two AdapterCutters: per read a dict name -> statistics
            cutter = rj["trim"]["modifiers"]["AdapterCutter"]
            ads = [[adapter_stats(st) for st in (d or {}).values()] for d in cutter["adapters"]]
        merged = None
        if merge is not None:
            mtext = ""
            if merge.get("output", True) and os.path.exists(p["merged.fq"]):
                with open(p["merged.fq"], "r", newline="") as fh:
                    mtext = fh.read()
            merged = {"out": mtext, "records_filtered": rj["trim"]["filters"]["MergedReadFilter"]["records_filtered"]}
        return {"ops": ops_stats(rj), "corrected": corrected, "merged": merged, "out1": outs[0], "out2": outs[1], "records": rj["record_counts"].get("0", 0),
                "with_adapters": cutter["records_with_adapters"], "bp_in": rj["bp_counts"].get("0", [0, 0]),
                "bp_out": rj["trim"]["formatters"]["bp_written"], "adapters": ads}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    build_ref.build()
    ref_loader.load_package()
    rng = np.random.default_rng(9101)
    cases = []

    def add(label, recs, error_rate=0.1, edit=None, extra=(), read_ops=None, adapter_args=None, times=1, mismatch_action=None, merge=None):
        if mismatch_action:
            extra = list(extra) + ["--correct-mismatches", mismatch_action]
        t1, t2 = fastq(recs[0]), fastq(recs[1])
        if edit:
            t1, t2 = edit(t1, t2)
        res = run_reference(t1, t2, error_rate, extra, adapter_args, merge)
        print(label, {k: (v if not isinstance(v, (str, list)) else (len(v) if isinstance(v, str) else v if len(v) < 3 else len(v)))
                      for k, v in res.items()})
        cases.append({"label": label, "text1": t1, "text2": t2, "error_rate": error_rate, "read_ops": read_ops or {},
                      "mode": "insert" if adapter_args is None else "adapter", "times": times,
                      "mismatch_action": mismatch_action, "merge": merge, "result": res})

    add("pe150", make_pairs(rng, 500, seed=11))
    add("pe150_suffix_names", make_pairs(rng, 300, suffix=True, seed=12))
    add("ragged_lower", make_pairs(rng, 500, ragged=True, lower=0.05, seed=13))
    add("default_rates", make_pairs(rng, 300, ragged=True, seed=14), error_rate=None)
    add("pe100_e02", make_pairs(rng, 300, L=100, seed=15), error_rate=0.2)
    add("empty_files", ([], []))
    # --- the paper's benchmark options (paper/workflow/simulated.nf:143-150: -q 0 --trim-n -m 25) and friends -----------
    def lowq(recs):
        out = []
        for name, seq, name2, q in recs:
            L = len(seq)
            q = "".join(chr(33 + int(max(2, min(40, 38 - i * rng.uniform(0.0, 0.35) + rng.normal(0, 4))))) for i in range(L))
            if rng.random() < 0.2 and L > 6:
                k1, k2 = int(rng.integers(0, 3)), int(rng.integers(0, 5))
                seq = "N" * k1 + seq[k1:L - k2] + "N" * k2
            out.append((name, seq, name2, q))
        return out
    r = make_pairs(rng, 400, ragged=True, seed=21)
    add("ops_trimn_minlen", (lowq(r[0]), lowq(r[1])), extra=["--trim-n", "-m", "25"], read_ops=dict(trim_n=True, minimum_length=25))
    r = make_pairs(rng, 400, ragged=True, seed=22)
    add("ops_quality_cut_maxn", (lowq(r[0]), lowq(r[1])), extra=["-q", "10,20", "-u", "3", "-U", "-4", "--max-n", "3", "-m", "30", "-M", "140"],
        read_ops=dict(quality_cutoff=[10, 20], cut=[3], cut2=[-4], max_n=3, minimum_length=30, maximum_length=140))
    r = make_pairs(rng, 300, seed=23)
    add("ops_discard_untrimmed", (lowq(r[0]), lowq(r[1])), extra=["--discard-untrimmed", "--trim-n"], read_ops=dict(discard_untrimmed=True, trim_n=True))
    # --- the command's default paired-end mode: --aligner adapter, independent adapter cutters per read ------------------
    FRONT5, SMALL3 = "GTTCAGAGTTCTACAGTCCGACGATC", "TGGAATTCTCGGGTGCCAAGG"
    r = make_pairs(rng, 400, ragged=True, lower=0.03, seed=24)
    add("adapter_mode_basic", r, adapter_args=["-a", A1, "-A", A2])
    r = make_pairs(rng, 400, ragged=True, seed=25)
    add("adapter_mode_panel_times2_ops", (lowq(r[0]), lowq(r[1])),
        adapter_args=["-a", A1, "-a", SMALL3, "-g", FRONT5, "-A", A2, "-G", FRONT5, "-n", "2"], times=2,
        extra=["-q", "10", "--trim-n", "-m", "20"], read_ops=dict(quality_cutoff=[10], trim_n=True, minimum_length=20))
    r = make_pairs(rng, 300, seed=26)
    add("adapter_mode_read1_only", r, adapter_args=["-a", A1], extra=["--discard-untrimmed"], read_ops=dict(discard_untrimmed=True, legacy_first=True))

    # --- --correct-mismatches (ErrorCorrectorMixin, modifiers.py:201-357): equal-length reads with mismatching overlaps ----
    def noisy(recs):
        out = []
        for name, seq, name2, q in recs:
            L = len(seq)
            q = "".join(chr(33 + int(max(2, min(40, rng.normal(30, 8))))) for _ in range(L))
            seq = "".join(("N" if rng.random() < 0.01 else c) for c in seq)
            out.append((name, seq, name2, q))
        return out
    for act, seed in (("liberal", 27), ("conservative", 28), ("N", 29)):
        r1_, r2_ = synth.synth_pe(300, 150, seed=seed, device="cpu", sub=0.03)
        recs = make_pairs(rng, 300, seed=seed)
        recs = ([(n_, bytes(r1_[i].numpy()).decode(), n2_, q_) for i, (n_, s_, n2_, q_) in enumerate(recs[0])],
                [(n_, bytes(r2_[i].numpy()).decode(), n2_, q_) for i, (n_, s_, n2_, q_) in enumerate(recs[1])])
        add("correct_" + act.lower(), (noisy(recs[0]), noisy(recs[1])), mismatch_action=act,
            extra=["--trim-n", "-m", "20"], read_ops=dict(trim_n=True, minimum_length=20))

    # --- improper pairing / malformed input --------------------------------------------------------------------
    def drop_last(which):
        def f(t1, t2):
            cut = lambda t: "\n".join(t.split("\n")[:-5]) + "\n"
            return (cut(t1), t2) if which == 1 else (t1, cut(t2))
        return f
    add("err_more_in_2", make_pairs(rng, 6, seed=16), edit=drop_last(1))
    add("err_more_in_1", make_pairs(rng, 6, seed=17), edit=drop_last(2))
    add("err_names", make_pairs(rng, 6, seed=18), edit=lambda t1, t2: (t1, t2.replace("@pair3", "@other3")))
    add("err_format_file2", make_pairs(rng, 6, seed=19), edit=lambda t1, t2: (t1, t2.replace("\n+\n", "\n-\n", 3).replace("\n-\n", "\n+\n", 2)))
    add("err_truncated_file1", make_pairs(rng, 6, seed=20), edit=lambda t1, t2: ("\n".join(t1.split("\n")[:-3]) + "\n", t2))

    # --- --merge-overlapping --merged-output (MergeOverlapping modifiers.py:864-931 + MergedReadFilter): added last so that the
    # cases above keep their random draws ---------------------------------------------------------------------------------
    add("merge_insert", make_pairs(rng, 400, seed=41), merge={})
    r = make_pairs(rng, 400, ragged=True, lower=0.03, seed=42)
    add("merge_insert_ragged_ops", (lowq(r[0]), lowq(r[1])), extra=["--trim-n", "-m", "25"], read_ops=dict(trim_n=True, minimum_length=25), merge={})
    add("merge_adapter_mode", make_pairs(rng, 300, ragged=True, seed=43), adapter_args=["-a", A1, "-A", A2], merge={"min_overlap": 0.5})
    add("merge_abs_overlap_e01", make_pairs(rng, 300, L=100, seed=44), merge={"min_overlap": 30, "error_rate": 0.1})
    add("merge_discarded", make_pairs(rng, 200, seed=45), merge={"output": False}, extra=["-m", "40"], read_ops=dict(minimum_length=40))
    for act, seed in (("liberal", 46), ("conservative", 47), ("N", 48)):
        r1_, r2_ = synth.synth_pe(300, 150, seed=seed, device="cpu", sub=0.03)
        recs = make_pairs(rng, 300, seed=seed)
        recs = ([(n_, bytes(r1_[i].numpy()).decode(), n2_, q_) for i, (n_, s_, n2_, q_) in enumerate(recs[0])],
                [(n_, bytes(r2_[i].numpy()).decode(), n2_, q_) for i, (n_, s_, n2_, q_) in enumerate(recs[1])])
        add("merge_correct_" + act.lower(), (noisy(recs[0]), noisy(recs[1])), mismatch_action=act, merge={"min_overlap": 0.6},
            extra=["--trim-n"], read_ops=dict(trim_n=True))
    r1_, r2_ = synth.synth_pe(300, 150, seed=49, device="cpu", sub=0.03)
    recs = make_pairs(rng, 300, seed=49)
    recs = ([(n_, bytes(r1_[i].numpy()).decode(), n2_, q_) for i, (n_, s_, n2_, q_) in enumerate(recs[0])],
            [(n_, bytes(r2_[i].numpy()).decode(), n2_, q_) for i, (n_, s_, n2_, q_) in enumerate(recs[1])])
    add("merge_adapter_mode_correct", (noisy(recs[0]), noisy(recs[1])), adapter_args=["-a", A1, "-A", A2], mismatch_action="liberal",
        merge={"min_overlap": 0.3, "error_rate": 0.25})

    # --- --pair-filter both: a pair is discarded only if both reads meet a filter's criterion (PairedWrapper min_affected 2) ----
    r = make_pairs(rng, 400, ragged=True, seed=50)
    add("pair_filter_both", (lowq(r[0]), lowq(r[1])), extra=["--pair-filter", "both", "--trim-n", "-m", "60", "--max-n", "1"],
        read_ops=dict(pair_filter="both", trim_n=True, minimum_length=60, max_n=1))
    r = make_pairs(rng, 300, ragged=True, seed=51)
    add("pair_filter_both_adapter_mode_merge", (lowq(r[0]), lowq(r[1])), adapter_args=["-a", A1, "-A", A2],
        extra=["--pair-filter", "both", "-m", "80", "--discard-untrimmed"], read_ops=dict(pair_filter="both", minimum_length=80, discard_untrimmed=True),
        merge={"min_overlap": 0.7})

    path = os.path.join(HERE, "fastq_trim_pe.json.gz")
    with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as fh:
        fh.write(json.dumps(cases, separators=(",", ":")).encode("ascii"))
    print("fastq_trim_pe.json.gz", os.path.getsize(path), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
