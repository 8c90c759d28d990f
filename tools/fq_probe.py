import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from atropos_b200 import synth, fastq
from atropos_b200.adapters import Adapter, BACK
n, L = 10_000_000, 150
dev = torch.device("cuda", 0)
reads = synth.synth_se(n, L, seed=synth.seed_for(2), device=dev).cpu().numpy()
text_np = synth.fastq_text(reads)
text = torch.empty(text_np.size, dtype=torch.uint8, pin_memory=True); text.numpy()[:] = text_np
out = torch.empty(text_np.size, dtype=torch.uint8, pin_memory=True)
ad = Adapter(synth.TRUSEQ_R1, BACK, max_error_rate=0.1, min_overlap=3)
# raw copy speeds
d = torch.empty(text_np.size, dtype=torch.uint8, device=dev)
for name, fn in (("h2d", lambda: d.copy_(text, non_blocking=True)), ("d2h", lambda: out.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (text_np.size / (time.perf_counter() - t0) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(text, non_blocking=True)
d2 = torch.empty_like(d)
with torch.cuda.stream(s2): out.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print("bidirectional: %.1f GB/s each" % (text_np.size / (time.perf_counter() - t0) / 1e9))
del d, d2
for chunk_mb in (16, 64, 256):
    tr = fastq.FastqTrimmer([ad], times=1, max_len=L, chunk_bytes=chunk_mb << 20)
    tr.trim(text.numpy(), out=out.numpy())
    t0 = time.perf_counter()
    for _ in range(3): tr.trim(text.numpy(), out=out.numpy())
    dt = (time.perf_counter() - t0) / 3
    print("chunk %d MB: %.1f ms, %.1f M reads/s" % (chunk_mb, dt * 1e3, n / dt / 1e6), flush=True)
