"""Rate of MergeOverlapping on the GPU (atr_merge_overlap_batch_host -> k_merge_overlap) next to the reference's
per-pair work on one host core (the reference's compiled Aligner from oracle/_ref: Aligner(rc(read 2), rate,
SEMIGLOBAL).locate(read 1), commands/trim/modifiers.py:886-895). Prints one JSON line per read length.

    python tests/merge_probe.py [--pairs 1000000] [--cpu-pairs 3000]

Lives under tests/ because it runs the reference (oracle/_ref) as the checker and the CPU baseline.
Pairs: both reads of length L from one fragment of length F ~ U[L, 3L] (half of the pairs overlap), 1 % substitutions.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from atropos_b200 import _abi, engine  # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def pairs(n, L, seed):
    rng = np.random.default_rng(seed)
    F = rng.integers(L, 3 * L + 1, size=n)
    r1 = np.empty((n, L), dtype=np.uint8)
    r2 = np.empty((n, L), dtype=np.uint8)
    for c0 in range(0, n, 200_000):
        c1 = min(n, c0 + 200_000)
        frag = rng.integers(0, 4, size=(c1 - c0, 3 * L), dtype=np.uint8)
        a = frag[:, :L].copy()
        col = (F[c0:c1, None] - 1 - np.arange(L)[None, :]).astype(np.int32)
        b = 3 - np.take_along_axis(frag, col, axis=1)
        for x in (a, b):                                        # substitutions
            hit = rng.random(x.shape) < 0.01
            x[hit] = (x[hit] + rng.integers(1, 4, size=int(hit.sum()), dtype=np.uint8)) & 3
        r1[c0:c1], r2[c0:c1] = ACGT[a], ACGT[b]
    return r1, r2, F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--cpu-pairs", type=int, default=3000)
    ap.add_argument("--rate", type=float, default=0.2)         # the command's default --merge-error-rate (trim/cli.py:690-691)
    ap.add_argument("--min-overlap", type=float, default=0.9)
    args = ap.parse_args()
    ctx = engine.default_context(0)
    for L in (150, 300):
        n = args.pairs if L == 150 else args.pairs // 4
        r1, r2, F = pairs(n, L, 77 + L)
        offs = engine.fixed_length_offsets(n, L)
        a1, a2 = r1.reshape(-1), r2.reshape(-1)
        ctx.merge_overlap_host(a1, offs, a2, offs, args.min_overlap, args.rate)                  # warm-up (allocations)
        t0 = time.perf_counter()
        recs = ctx.merge_overlap_host(a1, offs, a2, offs, args.min_overlap, args.rate)
        e2e_ms = (time.perf_counter() - t0) * 1e3
        ctx.set_profiling(True)
        ctx.merge_overlap_host(a1, offs, a2, offs, args.min_overlap, args.rate)
        k_ms = ctx.last_kernel_ms()
        ctx.set_profiling(False)
        line = {"what": "MergeOverlapping 2x%d, rate %.2f, min_overlap %.2f" % (L, args.rate, args.min_overlap), "pairs": n,
                "kernel_ms": k_ms, "kernel_M_pairs_per_s": n / k_ms / 1e3, "e2e_ms": e2e_ms, "e2e_M_pairs_per_s": n / e2e_ms / 1e3,
                "merged_fraction": float((recs["status"] == _abi.ATR_ST_MATCH).mean()),
                "overlapping_fraction": float((F < 2 * L).mean())}
        try:                                                    # the reference's compiled aligner, one core
            from oracle import ref_loader
            Aligner = ref_loader.load_native().Aligner
            comp = bytes.maketrans(b"ACGT", b"TGCA")
            m = min(args.cpu_pairs, n)
            s1 = [r1[i].tobytes().decode() for i in range(m)]
            s2 = [r2[i].tobytes().translate(comp)[::-1].decode() for i in range(m)]
            t0 = time.perf_counter()
            got = [Aligner(b, args.rate, 15).locate(a) for a, b in zip(s1, s2)]
            cpu_s = time.perf_counter() - t0
            same = all((g is None and int(r["matches"]) == 0) or (g is not None and tuple(int(r[f]) for f in (
                "r2_start", "r2_stop", "r1_start", "r1_stop", "matches", "errors")) == g) for g, r in zip(got, recs[:m]))
            line.update(cpu_pairs=m, cpu_M_pairs_per_s_one_core=m / cpu_s / 1e6, cpu_sample_identical=bool(same))
        except Exception as e:                                  # oracle/_ref not built here
            line["cpu"] = "unavailable: %r" % (e,)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
