#!/usr/bin/env python
"""Run the reference's OWN `trim` command (baseline/_ref, unmodified) on one paired-end golden case of
tests/golden/fastq_trim_pe.json.gz with `atropos.align._align` served by atropos_b200 and the batched TrimPipeline binding
(atropos_b200/integration.py), and compare what it writes with what the unmodified reference wrote when the golden was made.

    python tests/run_reference_cli.py --case merge_insert [--sim] [--mode batched|percall]

Prints ``ATR_CLI {json}`` (the binding's counters, which outputs matched) and exits 0 iff every output is identical.
Test infrastructure -- the product never imports it.
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGE = os.path.join(ROOT, "baseline", "_ref")


def main(argv):
    case_label, sim, mode = "merge_insert", False, "batched"
    it = iter(argv)
    for a in it:
        if a == "--case":
            case_label = next(it)
        elif a == "--sim":
            sim = True
        elif a == "--mode":
            mode = next(it)
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.path.insert(0, STAGE)
    import golden_util
    case = [c for c in golden_util.load("fastq_trim_pe") if c["label"] == case_label][0]
    if sim:
        import simbackend
        simbackend.install()
    from atropos_b200 import integration, synth
    integration.install(batched=(mode == "batched"))
    from atropos.commands import get_command
    tmp = tempfile.mkdtemp(prefix="atrcli")
    try:
        p = {k: os.path.join(tmp, k) for k in ("in1.fq", "in2.fq", "out1.fq", "out2.fq", "merged.fq", "rep")}
        for key, text in (("in1.fq", case["text1"]), ("in2.fq", case["text2"])):
            with open(p[key], "w", newline="") as fh:
                fh.write(text)
        if case.get("mode") == "adapter":
            args = ["-a", synth.TRUSEQ_R1, "-A", synth.TRUSEQ_R2]
        else:
            args = ["--aligner", "insert", "-a", synth.TRUSEQ_R1, "-A", synth.TRUSEQ_R2]
        args += ["-pe1", p["in1.fq"], "-pe2", p["in2.fq"], "-o", p["out1.fq"], "-p", p["out2.fq"], "--no-default-adapters",
                 "--no-cache-adapters", "--quiet", "--report-file", p["rep"], "--report-formats", "json"]
        ops = case.get("read_ops") or {}
        if ops.get("trim_n"):
            args += ["--trim-n"]
        if ops.get("minimum_length"):
            args += ["-m", str(ops["minimum_length"])]
        if ops.get("discard_untrimmed"):
            args += ["--discard-untrimmed"]
        if ops.get("pair_filter") == "both":
            args += ["--pair-filter", "both"]
        if case.get("mismatch_action"):
            args += ["--correct-mismatches", case["mismatch_action"]]
        if case["error_rate"] is not None:
            args += ["-e", repr(case["error_rate"])]
        m = case.get("merge")
        if m is not None:
            args += ["--merge-overlapping"]
            if m.get("output", True):
                args += ["--merged-output", p["merged.fq"]]
            if "min_overlap" in m:
                args += ["--merge-min-overlap", repr(m["min_overlap"])]
            if "error_rate" in m:
                args += ["--merge-error-rate", repr(m["error_rate"])]
        rc, summary = get_command("trim").execute(args)
        same = {}
        res = case["result"]
        for key, gold in (("out1.fq", res["out1"]), ("out2.fq", res["out2"]),
                          ("merged.fq", (res.get("merged") or {}).get("out") if m is not None and m.get("output", True) else None)):
            if gold is None:
                continue
            with open(p[key], "r", newline="") as fh:
                same[key] = fh.read() == gold
        info = dict(integration.STATS)
        info.update(case=case_label, rc=rc, same=same, mode=mode, sim=sim)
        print("ATR_CLI " + json.dumps(info))
        return 0 if rc == 0 and same and all(same.values()) else 1
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
