#!/usr/bin/env python
"""bench_extra.py -- kernel-level measurements of the OTHER BASELINE configurations (parity-test cases, not the
graded bench line): cfg1 (SE100 13-mer), cfg3 (PE 2x150 insert aligner), cfg4 (8-adapter panel, one GPU's share).
Prints one JSON object per config. Inputs resident in HBM, CUDA events on the ctx stream.

    python bench_extra.py [--pairs N] [--reads N] [--steps K]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def time_steps(stream, fn, steps, warmup=2):
    import torch
    torch.cuda.synchronize()
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


def pack(ctx, reads_dev, L):
    import torch
    from atropos_b200 import engine
    n = reads_dev.shape[0]
    dev = reads_dev.device
    offs = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    codes = torch.empty(n * ((L + 7) // 8) + 8, dtype=torch.int32, device=dev)
    woff = torch.empty(n + 1, dtype=torch.int32, device=dev)
    lens = torch.empty(n, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()          # the ctx stream does not synchronise with torch's streams
    engine._lib.check(ctx._L.atr_pack_device(ctx.handle, reads_dev.data_ptr(), offs.data_ptr(), n, 0, codes.data_ptr(),
                                             woff.data_ptr(), lens.data_ptr()), ctx.handle)
    ctx.sync()
    return codes, woff, lens, offs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--pairs", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from atropos_b200 import _abi, engine, synth
    from atropos_b200.adapters import Adapter, BACK, FRONT, PREFIX
    from atropos_b200.align import InsertAligner
    from atropos_b200.modifiers import AdapterCutter
    dev = torch.device("cuda", 0)
    ctx = engine.default_context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    peak = 6538.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    # ---- cfg1: SE 100 bp, 13-mer ------------------------------------------------------------------
    n, L = args.reads, 100
    reads = synth.synth_se(n, L, synth.SHORT_ADAPTER, seed=synth.seed_for(1), device=dev)
    codes, woff, lens, offs = pack(ctx, reads, L)
    out = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    aset = Adapter(synth.SHORT_ADAPTER, BACK, 0.1, 3)._adapterset()
    ms = time_steps(stream, lambda: aset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out.data_ptr()), args.steps)
    b = (L + 1) // 2 + 4 + 16
    print(json.dumps({"config": "cfg1 SE100 13-mer", "reads": n, "ms": ms, "M_reads_per_s": n / ms / 1e3,
                      "algo_bytes_per_read": b, "hbm_frac": b * n / (ms * 1e-3) / 1e9 / peak}))
    del reads, codes, woff, lens, offs, out

    # ---- cfg4 (one GPU's share): SE150, 8-adapter panel, best of N + linked -----------------------------
    n, L = min(args.reads, 5_000_000), 150
    reads = synth.synth_se(n, L, synth.TRUSEQ_R1, seed=synth.seed_for(4), device=dev)
    codes, woff, lens, offs = pack(ctx, reads, L)
    out = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    specs = [(synth.TRUSEQ_R1, BACK), (synth.TRUSEQ_R2, BACK), ("TGGAATTCTCGGGTGCCAAGG", BACK),
             ("GTTCAGAGTTCTACAGTCCGACGATC", PREFIX), ("ACACTCTTTCCCTACACGACGCTCTTCCGATCT", PREFIX),
             ("AATGATACGGCGACCACCGA", FRONT), ("TGGAATTCTCGGGTGCCAAGG", BACK), ("AGATCGGAAGAGC", BACK)]
    cutter = AdapterCutter([Adapter(s, w) for s, w in specs])
    pset = cutter._adapterset()
    ms = time_steps(stream, lambda: pset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out.data_ptr()), args.steps)
    b = (L + 1) // 2 + 4 + 16
    print(json.dumps({"config": "cfg4 SE150 8-adapter panel (best-of-N on GPU)", "reads": n, "ms": ms,
                      "M_reads_per_s": n / ms / 1e3, "algo_bytes_per_read": b,
                      "hbm_frac": b * n / (ms * 1e-3) / 1e9 / peak}))
    del reads, codes, woff, lens, offs, out

    # ---- cfg3: PE 2x150 insert aligner ------------------------------------------------------------------
    n, L = args.pairs, 150
    r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(3), device=dev)
    c1, w1, l1, _ = pack(ctx, r1, L)
    c2, w2, l2, _ = pack(ctx, r2, L)
    del r1, r2
    ia = InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1)
    iset = ia._insertset(L)
    iout = torch.empty((n, 48), dtype=torch.uint8, device=dev)
    ms = time_steps(stream, lambda: iset.match_insert_device(c1.data_ptr(), w1.data_ptr(), l1.data_ptr(), c2.data_ptr(),
                                                              w2.data_ptr(), l2.data_ptr(), n, iout.data_ptr()), args.steps)
    b = 2 * ((L + 1) // 2) + 8 + 48
    res = iout.cpu().numpy().view(_abi.INSERT_DTYPE).reshape(-1)
    print(json.dumps({"config": "cfg3 PE 2x150 insert aligner (match_insert)", "pairs": n, "ms": ms,
                      "M_pairs_per_s": n / ms / 1e3, "algo_bytes_per_pair": b,
                      "hbm_frac": b * n / (ms * 1e-3) / 1e9 / peak,
                      "insert_match_fraction": float((res["insert"]["status"] == 1).mean())}))
    del c1, w1, l1, c2, w2, l2, iout

    # ---- cfg5 (one GPU's share of the shape): PE 2x300, error rate 0.15 ----------------------------------------------
    n5, L5 = min(args.pairs, 4_000_000), 300
    r1, r2 = synth.synth_pe(n5, L5, seed=synth.seed_for(5), device=dev, sub=0.02)
    c1, w1, l1, _ = pack(ctx, r1, L5)
    c2, w2, l2, _ = pack(ctx, r2, L5)
    del r1, r2
    ia5 = InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, max_insert_mismatch_frac=0.15, max_adapter_mismatch_frac=0.15)
    iset5 = ia5._insertset(L5)
    iout = torch.empty((n5, 48), dtype=torch.uint8, device=dev)
    ms = time_steps(stream, lambda: iset5.match_insert_device(c1.data_ptr(), w1.data_ptr(), l1.data_ptr(), c2.data_ptr(),
                                                               w2.data_ptr(), l2.data_ptr(), n5, iout.data_ptr()), args.steps)
    b = 2 * ((L5 + 1) // 2) + 8 + 48
    res = iout.cpu().numpy().view(_abi.INSERT_DTYPE).reshape(-1)
    print(json.dumps({"config": "cfg5 PE 2x300 err 0.15 insert aligner (match_insert)", "pairs": n5, "ms": ms,
                      "M_pairs_per_s": n5 / ms / 1e3, "algo_bytes_per_pair": b,
                      "hbm_frac": b * n5 / (ms * 1e-3) / 1e9 / peak,
                      "insert_match_fraction": float((res["insert"]["status"] == 1).mean())}))
    del c1, w1, l1, c2, w2, l2, iout

    # ---- cfg3 end to end: two FASTQ texts (pinned host) -> atr_trim_fastq_pe_host -> two trimmed FASTQ texts --------
    import time
    from atropos_b200 import fastq
    from atropos_b200.util import RandomMatchProbability
    n = min(args.pairs, 4_000_000)
    r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(3), device=dev)
    texts, outs = [], []
    for r in (r1, r2):
        t = synth.fastq_text(r)
        th = torch.empty(t.size, dtype=torch.uint8, pin_memory=True)
        th.numpy()[:] = t
        texts.append(th)
        outs.append(torch.empty(t.size, dtype=torch.uint8, pin_memory=True))
    del r1, r2
    rmp = RandomMatchProbability()
    kw = dict(max_error_rate=0.1, min_overlap=1, indel_cost=3, max_rmp=1e-6, match_probability=rmp)
    tr = fastq.FastqPairTrimmer(Adapter(synth.TRUSEQ_R1, BACK, **kw), Adapter(synth.TRUSEQ_R2, BACK, **kw),
                                InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, match_probability=rmp,
                                              max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1), max_len=L)
    run = lambda: tr.trim(texts[0].numpy(), texts[1].numpy(), out1=outs[0].numpy(), out2=outs[1].numpy())
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        o, st, _ = run()
    dt = (time.perf_counter() - t0) / 2
    print(json.dumps({"config": "cfg3 PE 2x150 FASTQ text -> trimmed FASTQ text (atr_trim_fastq_pe_host, --aligner insert)",
                      "pairs": n, "ms": dt * 1e3, "M_pairs_per_s": n / dt / 1e6,
                      "h2d_bytes": int(texts[0].numel() + texts[1].numel()), "d2h_bytes": int(o[0].size + o[1].size),
                      "insert_matches": int(st.insert_matches), "with_adapters": st.with_adapters}))

    # ... and the same with --merge-overlapping --merged-output (atr_trim_fastq_pe_merge_host: merge kernels on what trimming left)
    trm = fastq.FastqPairTrimmer(Adapter(synth.TRUSEQ_R1, BACK, **kw), Adapter(synth.TRUSEQ_R2, BACK, **kw),
                                 InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, match_probability=rmp,
                                               max_insert_mismatch_frac=0.1, max_adapter_mismatch_frac=0.1), max_len=L,
                                 merge_overlapping=True, merge_min_overlap=0.9, merge_error_rate=0.2)
    outm = torch.empty(texts[0].numel() + texts[1].numel() + 1, dtype=torch.uint8, pin_memory=True)
    runm = lambda: trm.trim(texts[0].numpy(), texts[1].numpy(), out1=outs[0].numpy(), out2=outs[1].numpy(), out_merged=outm.numpy())
    runm()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        om, stm, _ = runm()
    dtm = (time.perf_counter() - t0) / 2
    print(json.dumps({"config": "cfg3 PE 2x150 FASTQ text -> trimmed + merged FASTQ text (atr_trim_fastq_pe_merge_host, --aligner insert "
                                "--merge-overlapping --merged-output)",
                      "pairs": n, "ms": dtm * 1e3, "M_pairs_per_s": n / dtm / 1e6,
                      "h2d_bytes": int(texts[0].numel() + texts[1].numel()), "d2h_bytes": int(om[0].size + om[1].size + om[2].size),
                      "merged_pairs": int(stm.merged), "merged_bytes": int(om[2].size), "insert_matches": int(stm.insert_matches)}))
    del outm, om

    # ---- MergeOverlapping: pairs of the cfg3 shape after trimming, i.e. both reads cut to the fragment ---
    import numpy as np
    del texts, o
    for Lm, nm in ((150, 1_000_000), (300, 250_000)):
        r1, r2 = synth.synth_pe(nm, Lm, seed=synth.seed_for(3) + Lm, device=dev, short_frac=0.0)    # fragments of L..3L: no adapter in the reads
        a1, a2 = r1.cpu().numpy().reshape(-1), r2.cpu().numpy().reshape(-1)
        offs = engine.fixed_length_offsets(nm, Lm)
        ctx.merge_overlap_host(a1, offs, a2, offs, 0.9, 0.2)                                      # warm-up (allocations)
        ctx.set_profiling(True)
        recs = ctx.merge_overlap_host(a1, offs, a2, offs, 0.9, 0.2)
        k_ms = ctx.last_kernel_ms()
        ctx.set_profiling(False)
        t0 = time.perf_counter()
        ctx.merge_overlap_host(a1, offs, a2, offs, 0.9, 0.2)
        dt = time.perf_counter() - t0
        b = 2 * Lm + 16 + 16
        print(json.dumps({"config": "MergeOverlapping 2x%d, error rate 0.2, min_overlap 0.9 (atr_merge_overlap_batch_host)" % Lm,
                          "pairs": nm, "kernel_ms": k_ms, "kernel_M_pairs_per_s": nm / k_ms / 1e3, "host_call_ms": dt * 1e3,
                          "host_call_M_pairs_per_s": nm / dt / 1e6, "algo_bytes_per_pair": b,
                          "hbm_frac": b * nm / (k_ms * 1e-3) / 1e9 / peak, "cell_updates_per_s": nm / (k_ms * 1e-3) * Lm * Lm,
                          "merged_fraction": float((recs["status"] == _abi.ATR_ST_MATCH).mean())}))


def reference_cli_backends(n_reads=500_000):
    """The reference's OWN command line (baseline/_ref, unmodified) on the same FASTQ file with its three aligner backends
    (atropos_b200.integration --aligner-backend): what the drop-in buys while every record still goes through the
    reference's per-record Python (reader, modifiers, filters, writers) -- and why the FASTQ-text entry points exist."""
    import subprocess
    import tempfile
    import time
    from atropos_b200 import synth
    stage = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(stage, "atropos")):
        print(json.dumps({"config": "reference CLI backends", "unavailable": "baseline/_ref not staged"}))
        return
    tmp = tempfile.mkdtemp(prefix="atrcli")
    reads = synth.synth_se(n_reads, 150, seed=synth.seed_for(2), device="cpu").numpy()
    with open(os.path.join(tmp, "in.fq"), "wb") as fh:
        fh.write(synth.fastq_text(reads).tobytes())
    out = {"config": "reference command line `atropos trim -a TRUSEQ -e 0.1 -se in.fq -o out.fq`, %d reads, one process, per backend" % n_reads}
    texts = {}
    for backend in ("cython", "gpu-per-call", "gpu"):
        dst = os.path.join(tmp, "out_%s.fq" % backend)
        cmd = [sys.executable, "-m", "atropos_b200.integration", "--aligner-backend", backend, "trim", "-a", synth.TRUSEQ_R1, "-e", "0.1",
               "--no-default-adapters", "--no-cache-adapters", "--quiet", "--report-file", os.devnull, "-se", os.path.join(tmp, "in.fq"), "-o", dst]
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, stage, os.environ.get("PYTHONPATH", "")]))
        t0 = time.perf_counter()
        p = subprocess.run(cmd, env=env, cwd=stage, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            out[backend] = "failed: " + p.stdout[-300:]
            continue
        out[backend + "_M_reads_per_s"] = n_reads / dt / 1e6
        with open(dst, "rb") as fh:
            texts[backend] = fh.read()
    out["outputs_identical"] = len(texts) == 3 and len(set(texts.values())) == 1
    print(json.dumps(out))
    # paired-end, the reference's flagship mode: --aligner insert (+ merging), where the aligner dominates its per-record cost
    n_pairs = n_reads // 4
    r1, r2 = synth.synth_pe(n_pairs, 150, seed=synth.seed_for(3), device="cpu")
    for name, r in (("in1.fq", r1), ("in2.fq", r2)):
        with open(os.path.join(tmp, name), "wb") as fh:
            fh.write(synth.fastq_text(r.numpy()).tobytes())
    for label, extra in (("insert", []), ("insert + merge", ["--merge-overlapping", "--merged-output", "MERGED"])):
        out = {"config": "reference command line `atropos trim --aligner insert -a R1 -A R2 -e 0.1 -pe1 .. -pe2 .. -o .. -p ..%s`, %d pairs, "
                         "one process, per backend" % (" -R --merged-output .." if extra else "", n_pairs)}
        texts = {}
        for backend in ("cython", "gpu"):
            dst = [os.path.join(tmp, "o%d_%s.fq" % (k, backend)) for k in (1, 2, 3)]
            cmd = [sys.executable, "-m", "atropos_b200.integration", "--aligner-backend", backend, "trim", "--aligner", "insert",
                   "-a", synth.TRUSEQ_R1, "-A", synth.TRUSEQ_R2, "-e", "0.1", "--no-default-adapters", "--no-cache-adapters", "--quiet",
                   "--report-file", os.devnull, "-pe1", os.path.join(tmp, "in1.fq"), "-pe2", os.path.join(tmp, "in2.fq"),
                   "-o", dst[0], "-p", dst[1]] + [dst[2] if x == "MERGED" else x for x in extra]
            env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, stage, os.environ.get("PYTHONPATH", "")]))
            t0 = time.perf_counter()
            p = subprocess.run(cmd, env=env, cwd=stage, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                out[backend] = "failed: " + p.stdout[-300:]
                continue
            out[backend + "_M_pairs_per_s"] = n_pairs / dt / 1e6
            texts[backend] = b"".join(open(d, "rb").read() for d in dst if os.path.exists(d))
        out["outputs_identical"] = len(texts) == 2 and len(set(texts.values())) == 1
        print(json.dumps(out))
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    if "--cli-backends" in sys.argv:
        reference_cli_backends()
    else:
        main()
