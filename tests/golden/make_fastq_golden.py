#!/usr/bin/env python
"""Generate tests/golden/fastq_trim.json.gz from the REAL reference command line (build container only).

    python tests/golden/make_fastq_golden.py

Every case is a small FASTQ text that is run through the unmodified reference, exactly as a user would:

    atropos trim <adapter options> -se in.fq -o out.fq --report-formats json

(FastqReader io/_seqio.pyx:163-245 -> AdapterCutter commands/trim/modifiers.py:91-195 -> Adapter.trimmed
adapters/__init__.py:413-436 -> FastqFormat io/seqio.py:686-700). Stored per case: the input text, the adapter
options, the output text and the statistics of the report (records, records with adapters, bp in / out, and per
adapter lengths_front/back, errors_front/back, adjacent_bases), or the FormatError the reader raised.
The GPU path (atr_trim_fastq_host) and its CPU simulation must reproduce all of it byte for byte.
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

import fuzzgen  # noqa: E402
from oracle import build_ref, ref_loader  # noqa: E402

TRUSEQ1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"
SMALL3 = "TGGAATTCTCGGGTGCCAAGG"
FRONT5 = "GTTCAGAGTTCTACAGTCCGACGATC"
ANY = "AATGATACGGCGACCACCGA"
WHERE_OPT = {"back": "-a", "front": "-g", "anywhere": "-b"}


def quals(rng, n):
    return "".join(chr(int(c)) for c in rng.integers(33, 74, n))


def fastq(records, eol="\n", final_eol=True):
    out = []
    for name, seq, name2, q in records:
        out.append("@" + name + eol + seq + eol + "+" + name2 + eol + q + eol)
    text = "".join(out)
    if not final_eol and text.endswith(eol):
        text = text[:-len(eol)]
    return text


def make_records(rng, n, adapters, max_len=150, ragged=True, lower=0.05, with_front=True):
    recs = []
    for i in range(n):
        seq_a, where = adapters[int(rng.integers(0, len(adapters)))]
        L = int(rng.integers(0, max_len + 1)) if ragged and rng.random() < 0.5 else max_len
        body = fuzzgen.read_with_adapter(rng, seq_a, max_len, n_rate=0.01)
        if where != "back" and with_front and rng.random() < 0.7:
            body = fuzzgen.mutate(rng, seq_a, 0.03, 0.01, 0.01) + body
        r = rng.random()
        if r < 0.02:
            body = seq_a + body                      # adapter at position 0: the read trims to nothing
        body = body[:L]
        if rng.random() < lower:
            body = body.lower()
        name = "read%d" % i + (" 1:N:0:%d" % int(rng.integers(0, 99)) if rng.random() < 0.5 else "")
        name2 = name if rng.random() < 0.2 else ""
        recs.append((name, body, name2, quals(rng, len(body))))
    return recs


def ops_stats(rj, read):
    """bp_trimmed of the cutters / trimmers, records_filtered of the filters, records written"""
    t = rj["trim"]
    mods, flt = t["modifiers"], t.get("filters", {})
    bp = lambda name: [int(x or 0) for x in mods[name]["bp_trimmed"]] if name in mods else None
    nf = lambda name: flt[name]["records_filtered"] if name in flt else None
    return {"bp_cut": bp("UnconditionalCutter"), "bp_quality": bp("QualityTrimmer"), "bp_n_ends": bp("NEndTrimmer"),
            "bp_nextseq": bp("NextseqQualityTrimmer"),
            "too_short": nf("too_short"), "too_long": nf("too_long"), "too_many_n": nf("too_many_n"),
            "discarded_trimmed": nf("TrimmedFilter"), "discarded_untrimmed": nf("UntrimmedFilter"),
            "records_written": t["formatters"]["records_written"]}


def run_reference(text, adapters, times, error_rate, overlap, extra=()):
    from atropos.commands import get_command
    tmp = tempfile.mkdtemp(prefix="fqgold")
    try:
        inp, outp, rep = (os.path.join(tmp, f) for f in ("in.fq", "out.fq", "rep"))
        with open(inp, "w", newline="") as fh:      # newline="": write the text untranslated
            fh.write(text)
        args = []
        for seq, where in adapters:
            args += [WHERE_OPT.get(where, "-a"), seq]          # "linked": -a FRONT...BACK
        args += ["-se", inp, "-o", outp, "-n", str(times), "-e", repr(error_rate), "-O", str(overlap),
                 "--no-default-adapters", "--no-cache-adapters", "--quiet", "--report-file", rep,
                 "--report-formats", "json"] + list(extra)
        try:
            rc, summary = get_command("trim").execute(args)
        except Exception as exc:                       # pragma: no cover
            return {"exception": repr(exc)}
        if rc != 0:
            # the command logs the exception and returns 1; the message itself comes from the reader
            from atropos.io.seqio import FormatError
            from atropos.io._seqio import FastqReader
            try:
                with FastqReader(inp) as reader:
                    for _ in reader:
                        pass
            except FormatError as exc:
                return {"error": str(exc)}
            return {"exception": "trim returned %r but the reader raised nothing" % (rc,)}
        with open(outp, "r", newline="") as fh:
            out_text = fh.read()
        with open(rep + ".json" if os.path.exists(rep + ".json") else rep) as fh:
            rj = json.load(fh)
        cutter = rj["trim"]["modifiers"]["AdapterCutter"]
        ad_stats = []
        for name, st in cutter["adapters"][0].items():
            if st["where"]["name"] == "linked":        # LinkedAdapter.summarize (adapters/__init__.py:705-745)
                for part, where in (("front", "front"), ("back", "back")):
                    d = {"name": name, "linked": part, "sequence": st[part + "_sequence"], "where": where}
                    for key in ("lengths_front", "lengths_back"):
                        d[key] = st["%s_%s" % (part, key)]
                    for key in ("errors_front", "errors_back"):
                        e = st["%s_%s" % (part, key)]
                        cols = e.get("columns", [])
                        d[key] = {ln: {str(c): v for c, v in zip(cols, row) if v} for ln, row in e.get("rows", {}).items()}
                    ad_stats.append(d)
                continue
            # the command line groups the adapters by option (-a, then -b, then -g): this list is the order the
            # AdapterCutter tries them in, which decides ties (modifiers.py:107-122)
            d = {"name": name, "sequence": st["sequence"], "where": st["where"]["name"]}
            for key in ("lengths_front", "lengths_back", "adjacent_bases"):
                if key in st:
                    d[key] = st[key]
            for key in ("errors_front", "errors_back"):
                if key in st:
                    cols = st[key]["columns"]
                    d[key] = {ln: {str(c): v for c, v in zip(cols, row) if v} for ln, row in st[key]["rows"].items()}
            ad_stats.append(d)
        extra_stats = ops_stats(rj, 0)
        return {"ops": extra_stats,
                "out": out_text, "records": rj["record_counts"].get("0", 0), "with_adapters": cutter["records_with_adapters"][0],
                "bp_in": rj["bp_counts"].get("0", [0, 0])[0],
                "bp_out": rj["trim"]["formatters"]["bp_written"][0] if "formatters" in rj["trim"] else None,
                "adapters": ad_stats}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    build_ref.build()
    ref_loader.load_package()
    rng = np.random.default_rng(9001)
    cases = []

    def add(label, text, adapters, times=1, error_rate=0.1, overlap=3, extra=(), read_ops=None):
        res = run_reference(text, adapters, times, error_rate, overlap, extra)
        print(label, {k: (v if not isinstance(v, (str, list)) else len(v)) for k, v in res.items()})
        cases.append({"label": label, "text": text, "adapters": adapters, "times": times, "error_rate": error_rate,
                      "overlap": overlap, "read_ops": read_ops or {}, "result": res})

    one = [(TRUSEQ1, "back")]
    add("se150_truseq", fastq(make_records(rng, 700, one, ragged=False, lower=0.0)), one)
    add("ragged_lower_n", fastq(make_records(rng, 600, one, ragged=True, lower=0.1)), one)
    panel = [(TRUSEQ1, "back"), (SMALL3, "back"), (FRONT5, "front"), (ANY, "anywhere")]
    add("panel_times1", fastq(make_records(rng, 700, panel)), panel)
    add("panel_times3", fastq(make_records(rng, 700, panel)), panel, times=3)
    add("front_only_e02", fastq(make_records(rng, 300, [(FRONT5, "front")])), [(FRONT5, "front")], error_rate=0.2, overlap=5)
    add("crlf", fastq(make_records(rng, 200, one), eol="\r\n"), one)
    add("no_final_newline", fastq(make_records(rng, 50, one, ragged=False), final_eol=False), one)
    add("empty_file", "", one)
    add("single_empty_read", "@r\n\n+\n\n", one)
    # --- a linked adapter: anchored 5' adapter, then (only if it matched) the 3' adapter in the remainder ------------------
    def linked_records(n):
        recs = []
        for i in range(n):
            body = fuzzgen.read_with_adapter(rng, SMALL3, 150, n_rate=0.01)
            r = rng.random()
            if r < 0.6:
                body = fuzzgen.mutate(rng, FRONT5, 0.03, 0.01, 0.01) + body
            elif r < 0.7:
                body = FRONT5[int(rng.integers(1, 10)):] + body          # partial: an anchored adapter must not match
            body = body[:int(rng.integers(20, 151))]
            recs.append(("lk%d" % i, body, "", quals(rng, len(body))))
        return recs
    add("linked_adapter", fastq(linked_records(500)), [(FRONT5 + "..." + SMALL3, "linked")])
    add("linked_adapter_ops", fastq(linked_records(300)), [(FRONT5 + "..." + SMALL3, "linked")],
        extra=["--trim-n", "-m", "30", "--discard-untrimmed"], read_ops=dict(trim_n=True, minimum_length=30, discard_untrimmed=True))

    # --- the modifiers / filters around the adapter stage (default operation order) --------------------------------
    def lowq(recs, rng):                                   # qualities that decay towards the ends, some N ends
        out = []
        for name, seq, name2, q in recs:
            L = len(seq)
            q = "".join(chr(33 + int(max(2, min(40, 40 - abs(i - L * 0.4) * rng.uniform(0.2, 0.9) + rng.normal(0, 4))))) for i in range(L))
            if rng.random() < 0.2 and L > 6:
                k1, k2 = int(rng.integers(0, 4)), int(rng.integers(0, 5))
                seq = "N" * k1 + seq[k1:L - k2] + "N" * k2
            if rng.random() < 0.02:
                seq = "N" * L
            out.append((name, seq, name2, q))
        return out
    add("ops_quality_trimn_minlen", fastq(lowq(make_records(rng, 500, one, ragged=True), rng)), one,
        extra=["-q", "15,20", "--trim-n", "-m", "25"], read_ops=dict(quality_cutoff=[15, 20], trim_n=True, minimum_length=25))
    add("ops_cut_maxlen_maxn", fastq(lowq(make_records(rng, 500, one, ragged=True), rng)), one,
        extra=["-u", "5", "-u", "-3", "-M", "120", "--max-n", "2"], read_ops=dict(cut=[5, -3], maximum_length=120, max_n=2))
    add("ops_q_single_maxn_frac_discard_untrimmed", fastq(lowq(make_records(rng, 400, one, ragged=True), rng)), one,
        extra=["-q", "20", "--max-n", "0.05", "--discard-untrimmed"], read_ops=dict(quality_cutoff=[20], max_n=0.05, discard_untrimmed=True))
    def nextseq(recs):                                      # dark cycles: runs of high-quality G at the 3' end
        out = []
        for name, seq, name2, q in recs:
            if rng.random() < 0.5 and len(seq) > 20:
                g = min(int(rng.integers(1, 30)), len(seq))
                seq = seq[:len(seq) - g] + "G" * g
            out.append((name, seq, name2, q))
        return out
    add("ops_nextseq_quality", fastq(nextseq(lowq(make_records(rng, 400, one, ragged=True), rng))), one,
        extra=["--nextseq-trim", "20", "-q", "10", "-m", "15"], read_ops=dict(nextseq_trim=20, quality_cutoff=[10], minimum_length=15))
    add("ops_panel_times2_all", fastq(lowq(make_records(rng, 500, panel), rng)), panel, times=2,
        extra=["-u", "2", "-q", "10,10", "--trim-n", "-m", "20", "-M", "140", "--discard-trimmed"],
        read_ops=dict(cut=[2], quality_cutoff=[10, 10], trim_n=True, minimum_length=20, maximum_length=140, discard_trimmed=True))
    # --- malformed inputs: the reader's FormatErrors -------------------------------------------------
    good = make_records(rng, 8, one, ragged=False)
    t = fastq(good)
    lines = t.split("\n")
    bad = list(lines); bad[8] = "r2 without at"
    add("err_no_at", "\n".join(bad), one)
    bad = list(lines); bad[14] = "-"
    add("err_no_plus", "\n".join(bad), one)
    bad = list(lines); bad[6] = "+othername"
    add("err_name_mismatch", "\n".join(bad), one)
    bad = list(lines); bad[11] = bad[11][:-3]
    add("err_qual_len", "\n".join(bad), one)
    add("err_truncated", "\n".join(lines[:18]) + "\n", one)
    add("err_truncated_mid", "\n".join(lines[:17]) + "\n", one)
    # --- nothing is removed and the last line has no newline: the output is ONE BYTE LONGER than the input (appended last
    # so that the cases above keep their random streams) ----------------------------------------------------------
    plain = [("u%d" % i, fuzzgen.rand_seq(rng, 60), "", quals(rng, 60)) for i in range(12)]
    add("no_final_newline_untrimmed", fastq(plain, final_eol=False), one, overlap=20)

    path = os.path.join(HERE, "fastq_trim.json.gz")
    with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as fh:
        fh.write(json.dumps(cases, separators=(",", ":")).encode("ascii"))
    print("fastq_trim.json.gz", os.path.getsize(path), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
