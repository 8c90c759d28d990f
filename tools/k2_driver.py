#!/usr/bin/env python
"""Small driver for profiling the insert-aligner kernels: python tools/k2_driver.py [--len 150] [--rate 0.1] [--pairs N]"""
import argparse
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--rate", type=float, default=0.1)
    ap.add_argument("--pairs", type=int, default=4_000_000)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from atropos_b200 import engine, synth
    from atropos_b200.align import InsertAligner
    dev = torch.device("cuda", 0)
    ctx = engine.default_context(0)
    n, L = a.pairs, a.len
    r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(3 if L == 150 else 5), device=dev, sub=0.02 if L == 300 else 0.01)
    offs = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    packed = []
    for r in (r1, r2):
        codes = torch.empty(n * ((L + 7) // 8) + 8, dtype=torch.int32, device=dev)
        woff = torch.empty(n + 1, dtype=torch.int32, device=dev)
        lens = torch.empty(n, dtype=torch.int16, device=dev)
        torch.cuda.synchronize()
        engine._lib.check(ctx._L.atr_pack_device(ctx.handle, r.data_ptr(), offs.data_ptr(), n, 0, codes.data_ptr(), woff.data_ptr(), lens.data_ptr()), ctx.handle)
        ctx.sync()
        packed.append((codes, woff, lens))
    ia = InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, max_insert_mismatch_frac=a.rate, max_adapter_mismatch_frac=a.rate)
    iset = ia._insertset(L)
    out = torch.empty((n, 48), dtype=torch.uint8, device=dev)
    (c1, w1, l1), (c2, w2, l2) = packed
    run = lambda: iset.match_insert_device(c1.data_ptr(), w1.data_ptr(), l1.data_ptr(), c2.data_ptr(), w2.data_ptr(), l2.data_ptr(), n, out.data_ptr())
    run(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    ctx.sync()
    dt = (time.perf_counter() - t0) / a.steps
    print("2x%d rate %.2f: %.3f ms per %d pairs = %.1f M pairs/s" % (L, a.rate, dt * 1e3, n, n / dt / 1e6))


if __name__ == "__main__":
    main()
