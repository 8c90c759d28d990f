#!/usr/bin/env python
"""Per-adapter and whole-panel kernel times on the cfg-4 shape: python tools/panel_driver.py [--reads N]"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=5_000_000)
    a = ap.parse_args()
    import torch
    import bench
    from atropos_b200 import adapters as ad_mod, engine, synth
    from atropos_b200.modifiers import AdapterCutter
    dev = torch.device("cuda", 0)
    ctx = engine.default_context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
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

    def t(aset, steps=5):
        f = lambda: aset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out.data_ptr())
        f(); f(); ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            f()
        e1.record(stream); e1.synchronize()
        return e0.elapsed_time(e1) / steps
    tot = 0.0
    for s, w in bench.PANEL:
        ms = t(ad_mod.Adapter(s, getattr(ad_mod, w), max_error_rate=0.1, min_overlap=3)._adapterset())
        tot += ms
        aset1 = ad_mod.Adapter(s, getattr(ad_mod, w), max_error_rate=0.1, min_overlap=3)._adapterset()
        ctx.set_profiling(True)
        aset1.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out.data_ptr())
        ph, names = ctx.last_phase_ms(), ctx.last_phase_names()
        ctx.set_profiling(False)
        print("%-60s %-7s %.3f ms per %d reads  %s" % (s, w, ms, n, dict(zip(names, [round(x, 3) for x in ph]))))
    cutter = AdapterCutter([ad_mod.Adapter(s, getattr(ad_mod, w), max_error_rate=0.1, min_overlap=3) for s, w in bench.PANEL])
    ms = t(cutter._adapterset())
    print("sum of single passes %.3f ms; panel %.3f ms = %.1f M reads/s" % (tot, ms, n / ms / 1e3))


if __name__ == "__main__":
    main()
