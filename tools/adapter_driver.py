#!/usr/bin/env python
"""One adapter of the cfg-4 panel against the cfg-4 reads, a few device calls (an ncu target for a single adapter's kernels):

    python tools/adapter_driver.py --adapter AGATCGG... --where BACK [--reads 5000000] [--steps 1]
"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--adapter", default="AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT")
    ap.add_argument("--where", default="BACK")
    ap.add_argument("--reads", type=int, default=5_000_000)
    ap.add_argument("--steps", type=int, default=1)
    a = ap.parse_args()
    import torch
    import bench
    from atropos_b200 import adapters as ad_mod, engine, synth
    dev = torch.device("cuda", 0)
    ctx = engine.default_context(0)
    n, L = a.reads, 150
    reads = synth.synth_se(n, L, bench.ADAPTER, seed=synth.seed_for(4), device=dev)
    offs = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    codes = torch.empty(n * 19 + 8, dtype=torch.int32, device=dev)
    woff = torch.empty(n + 1, dtype=torch.int32, device=dev)
    lens = torch.empty(n, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    engine._lib.check(ctx._L.atr_pack_device(ctx.handle, reads.data_ptr(), offs.data_ptr(), n, 1, codes.data_ptr(), woff.data_ptr(), lens.data_ptr()), ctx.handle)
    ctx.sync()
    out = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    aset = ad_mod.Adapter(a.adapter, getattr(ad_mod, a.where), max_error_rate=0.1, min_overlap=3)._adapterset()
    ctx.set_profiling(True)
    for _ in range(a.steps + 1):
        aset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out.data_ptr())
        ctx.sync()
    print(dict(zip(ctx.last_phase_names(), ctx.last_phase_ms())))
    import numpy as np
    from atropos_b200 import _abi
    res = out.cpu().numpy().view(_abi.MATCH_DTYPE).reshape(-1)
    print("match fraction", float((res["status"] == _abi.ATR_ST_MATCH).mean()))


if __name__ == "__main__":
    main()
