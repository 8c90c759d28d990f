"""Per-adapter kernel time of the cfg4 panel (one adapter set each) -- where does the panel's step go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from atropos_b200 import engine, synth
from atropos_b200.adapters import Adapter, BACK, FRONT, PREFIX
from bench_extra import pack, time_steps
dev = torch.device("cuda", 0)
ctx = engine.default_context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
n, L = 5_000_000, 150
reads = synth.synth_se(n, L, synth.TRUSEQ_R1, seed=synth.seed_for(4), device=dev)
codes, woff, lens, offs = pack(ctx, reads, L)
out = torch.empty((n, 16), dtype=torch.uint8, device=dev)
specs = [(synth.TRUSEQ_R1, BACK), (synth.TRUSEQ_R2, BACK), ("TGGAATTCTCGGGTGCCAAGG", BACK),
         ("GTTCAGAGTTCTACAGTCCGACGATC", PREFIX), ("ACACTCTTTCCCTACACGACGCTCTTCCGATCT", PREFIX),
         ("AATGATACGGCGACCACCGA", FRONT), ("TGGAATTCTCGGGTGCCAAGG", BACK), ("AGATCGGAAGAGC", BACK)]
ctx.set_profiling(True)
for s, w in specs:
    aset = Adapter(s, w)._adapterset()
    fn = lambda: aset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out.data_ptr())
    ms = time_steps(stream, fn, 5)
    fn(); ph = ctx.last_phase_ms()
    print("%-60s where=%2d m=%2d  %.3f ms  phases %s" % (s, w, len(s), ms, ["%.3f" % p for p in ph]), flush=True)
