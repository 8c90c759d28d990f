"""K1 as a warp-per-read wavefront (the north star's wording) against the shipped thread-per-read funnel: a measured A/B.

The warp wavefront exists as k_merge_warp (atr_merge_api.cuh: lanes own rows, one SHFL.UP per step). It aligns read 1
against rc(read 2) for every pair, so handing it rc(adapter) as every pair's read 2 makes it compute exactly the
34 x 150 adapter-vs-read matrix of cfg 2 (every cell, semiglobal flags, rate 0.1) -- the DP size and dataflow of a warp-per-read K1. Printed:
its kernel rate, next to the funnel's (atr_locate_batch_device on the same reads) and the register DP over every cell
(ATR_DISABLE_FUSED=1 in a second process).

    python tools/k1_warp_probe.py [--reads 4000000]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    args = ap.parse_args()
    import torch
    from atropos_b200 import engine, synth
    from atropos_b200.adapters import Adapter, BACK
    n, L = args.reads, 150
    reads = synth.synth_se(n, L, seed=synth.seed_for(2), device="cuda").cpu().numpy()
    offs = engine.fixed_length_offsets(n, L)
    ctx = engine.default_context(0)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rc = synth.TRUSEQ_R1.encode().translate(comp)[::-1]
    m = len(rc)
    a2 = np.tile(np.frombuffer(rc, dtype=np.uint8), n)
    offs2 = engine.fixed_length_offsets(n, m)
    a1 = reads.reshape(-1)
    ctx.merge_overlap_host(a1, offs, a2, offs2, 3, 0.1)                      # warm-up (allocations)
    ctx.set_profiling(True)
    ctx.merge_overlap_host(a1, offs, a2, offs2, 3, 0.1)
    warp_ms = ctx.last_kernel_ms()
    ctx.set_profiling(False)
    ad = Adapter(synth.TRUSEQ_R1, BACK, max_error_rate=0.1, min_overlap=3)
    ad.match_to_batch((a1, offs))
    line = {"what": "K1 mapping A/B, cfg 2 reads (34-nt adapter x 150-nt read)", "reads": n,
            "warp_wavefront_kernel": "k_merge_warp (one warp = one read, half-warp mode: 3 rows per lane, two reads per warp)",
            "warp_wavefront_ms": warp_ms, "warp_wavefront_M_reads_per_s": n / warp_ms / 1e3,
            "warp_wavefront_T_cells_per_s": n * m * L / warp_ms / 1e9}
    # the shipped funnel and the every-cell register DP on the same reads, device-resident, via the bench rig's entry point
    import subprocess
    for label, env in (("funnel", {}), ("register_dp_every_cell", {"ATR_DISABLE_FUSED": "1"})):
        code = ("import sys, json; sys.path.insert(0, %r)\n"
                "import torch\n"
                "from atropos_b200 import engine, synth, _lib\n"
                "from atropos_b200.adapters import Adapter, BACK\n"
                "n, L = %d, 150\n"
                "reads = synth.synth_se(n, L, seed=synth.seed_for(2), device='cuda')\n"
                "ctx = engine.default_context(0)\n"
                "ad = Adapter(synth.TRUSEQ_R1, BACK, max_error_rate=0.1, min_overlap=3)\n"
                "offs = torch.arange(n + 1, dtype=torch.int64, device='cuda') * L\n"
                "codes = torch.empty(n * ((L + 7) // 8) + 8, dtype=torch.int32, device='cuda')\n"
                "woff = torch.empty(n + 1, dtype=torch.int32, device='cuda')\n"
                "lens = torch.empty(n, dtype=torch.int16, device='cuda')\n"
                "torch.cuda.synchronize()\n"
                "_lib.check(ctx._L.atr_pack_device(ctx.handle, reads.data_ptr(), offs.data_ptr(), n, 1, codes.data_ptr(), woff.data_ptr(), lens.data_ptr()), ctx.handle)\n"
                "ctx.sync()\n"
                "out = torch.empty((n, 16), dtype=torch.uint8, device='cuda')\n"
                "aset = ad._adapterset()\n"
                "f = lambda: aset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out.data_ptr())\n"
                "f(); ctx.sync()\n"
                "st = torch.cuda.ExternalStream(ctx.stream)\n"
                "e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)\n"
                "e0.record(st)\n"
                "for _ in range(5): f()\n"
                "e1.record(st)\n"
                "e1.synchronize()\n"
                "print(json.dumps({'ms': e0.elapsed_time(e1) / 5}))\n") % (ROOT, n)
        try:
            outp = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
            ms = json.loads(outp.stdout.strip().splitlines()[-1])["ms"]
            line[label + "_ms"] = ms
            line[label + "_M_reads_per_s"] = n / ms / 1e3
        except Exception as exc:                                            # API names differ: report, do not die
            line[label] = "unavailable: %r %s" % (exc, outp.stderr[-300:] if "outp" in dir() else "")
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
