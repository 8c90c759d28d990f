"""Deterministic synthetic read generator for the benchmark configurations.

Written with torch ops so the same code generates small batches on the CPU (tests) and 10 M-read
batches directly in HBM on the GPU (bench.py). Bases i.i.d. uniform ACGT; fragment length "short"
(read runs into the adapter) with probability 0.40 ~ U[ceil(L/5), L-1], else U[L, 3L];
read = (fragment + adapter + random tail)[:L]; per-base substitution 0.01, at most one 1-base
insertion / deletion per read with probability L*1e-4 each, N with probability 1e-3.
Seeds: 20261017 + 1000*config + shard.
"""
import torch

TRUSEQ_R1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"                                  # doc/guide.rst:1172
TRUSEQ_R2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT"          # doc/guide.rst:1179
SHORT_ADAPTER = "AGATCGGAAGAGC"
BASE_SEED = 20261017

_ACGT = torch.tensor([65, 67, 71, 84], dtype=torch.uint8)
_COMP_IDX = torch.tensor([3, 2, 1, 0], dtype=torch.int64)        # A<->T, C<->G on 0..3 codes


def seed_for(config, shard=0):
    return BASE_SEED + 1000 * int(config) + int(shard)


def _noise(reads_idx, L, g, device, sub, indel, n_rate):
    """reads_idx: int64 [n, W] of base codes 0..3 (W >= L+1). Returns uint8 ASCII [n, L]."""
    n, W = reads_idx.shape
    if sub > 0:
        m = torch.rand((n, W), generator=g, device=device) < sub
        rnd = torch.randint(0, 4, (n, W), generator=g, device=device)
        reads_idx = torch.where(m, rnd, reads_idx)
    if indel > 0:
        ev = torch.rand(n, generator=g, device=device)
        pos = torch.randint(0, L, (n,), generator=g, device=device)
        has_del = (ev < indel * L)
        has_ins = (ev >= indel * L) & (ev < 2 * indel * L)
        col = torch.arange(W, device=device)[None, :]
        idx = col + (has_del[:, None] & (col >= pos[:, None])).long() - (has_ins[:, None] & (col > pos[:, None])).long()
        idx = idx.clamp_(0, W - 1)
        reads_idx = torch.gather(reads_idx, 1, idx)
        ins_base = torch.randint(0, 4, (n,), generator=g, device=device)
        rows = torch.nonzero(has_ins).squeeze(1)
        if rows.numel():
            reads_idx[rows, (pos[rows] + 1).clamp_(max=W - 1)] = ins_base[rows]
    out = _ACGT.to(device)[reads_idx[:, :L]]
    if n_rate > 0:
        nm = torch.rand((n, L), generator=g, device=device) < n_rate
        out = torch.where(nm, torch.tensor(78, dtype=torch.uint8, device=device), out)
    return out.contiguous()


def _encode(adapter, device):
    lut = {"A": 0, "C": 1, "G": 2, "T": 3}
    return torch.tensor([lut[c] for c in adapter], dtype=torch.int64, device=device)


def _fragment_lengths(n, L, g, device, short_frac):
    short = torch.rand(n, generator=g, device=device) < short_frac
    lo = -(-L // 5)
    f_short = torch.randint(lo, L, (n,), generator=g, device=device)
    f_long = torch.randint(L, 3 * L + 1, (n,), generator=g, device=device)
    return torch.where(short, f_short, f_long)


def synth_se(n, L, adapter=TRUSEQ_R1, seed=BASE_SEED, device="cpu", short_frac=0.40, sub=0.01, indel=1e-4,
             n_rate=1e-3, chunk=1 << 20):
    """Single-end reads: uint8 ASCII tensor [n, L] on `device`."""
    device = torch.device(device)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    a = _encode(adapter, device)
    W = L + 8
    outs = []
    for c0 in range(0, n, chunk):
        cn = min(chunk, n - c0)
        idx = torch.randint(0, 4, (cn, W), generator=g, device=device)
        frag = _fragment_lengths(cn, L, g, device, short_frac)
        rel = torch.arange(W, device=device)[None, :] - frag[:, None]
        inside = (rel >= 0) & (rel < a.numel())
        idx = torch.where(inside, a[rel.clamp(0, a.numel() - 1)], idx)
        outs.append(_noise(idx, L, g, device, sub, indel, n_rate))
    return torch.cat(outs, 0) if len(outs) != 1 else outs[0]


def synth_pe(n, L, adapter1=TRUSEQ_R1, adapter2=TRUSEQ_R2, seed=BASE_SEED, device="cpu", short_frac=0.40, sub=0.01,
             indel=1e-4, n_rate=1e-3, chunk=1 << 19):
    """Paired-end reads: two uint8 ASCII tensors [n, L]; R2 = (rc(fragment) + adapter2 + tail)[:L]."""
    device = torch.device(device)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    a1, a2 = _encode(adapter1, device), _encode(adapter2, device)
    W = L + 8
    comp = _COMP_IDX.to(device)
    o1, o2 = [], []
    for c0 in range(0, n, chunk):
        cn = min(chunk, n - c0)
        F = torch.randint(0, 4, (cn, 3 * L + 1), generator=g, device=device)
        frag = _fragment_lengths(cn, L, g, device, short_frac)
        col = torch.arange(W, device=device)[None, :]
        rel = col - frag[:, None]
        tail1 = torch.randint(0, 4, (cn, W), generator=g, device=device)
        tail2 = torch.randint(0, 4, (cn, W), generator=g, device=device)
        r1 = torch.gather(F, 1, col.expand(cn, W).clamp(max=3 * L))
        r2 = comp[torch.gather(F, 1, (frag[:, None] - 1 - col).clamp(0, 3 * L))]
        in1 = (rel >= 0) & (rel < a1.numel())
        in2 = (rel >= 0) & (rel < a2.numel())
        r1 = torch.where(rel < 0, r1, torch.where(in1, a1[rel.clamp(0, a1.numel() - 1)], tail1))
        r2 = torch.where(rel < 0, r2, torch.where(in2, a2[rel.clamp(0, a2.numel() - 1)], tail2))
        o1.append(_noise(r1, L, g, device, sub, indel, n_rate))
        o2.append(_noise(r2, L, g, device, sub, indel, n_rate))
    if len(o1) == 1:
        return o1[0], o2[0]
    return torch.cat(o1, 0), torch.cat(o2, 0)


def fastq_text(reads, name_prefix="r", qual_char=73):
    """reads: uint8 array/tensor [n, L] (ASCII) -> FASTQ text as a uint8 numpy array, records
    '@<prefix><9-digit index>\\n<read>\\n+\\n<qualities>\\n' (fixed-width names keep the construction vectorised)."""
    import numpy as np
    reads = reads.cpu().numpy() if hasattr(reads, "cpu") else np.asarray(reads)
    n, L = reads.shape
    p = np.frombuffer(name_prefix.encode("ascii"), dtype=np.uint8)
    H = 1 + len(p) + 9
    rec = np.empty((n, H + 1 + L + 1 + 2 + L + 1), dtype=np.uint8)
    rec[:, 0] = 64
    rec[:, 1:1 + len(p)] = p
    idx = np.arange(n, dtype=np.int64)
    for d in range(9):
        rec[:, H - 1 - d] = 48 + (idx // 10 ** d) % 10
    rec[:, H] = 10
    rec[:, H + 1:H + 1 + L] = reads
    rec[:, H + 1 + L] = 10
    rec[:, H + 2 + L] = 43
    rec[:, H + 3 + L] = 10
    rec[:, H + 4 + L:H + 4 + 2 * L] = qual_char
    rec[:, H + 4 + 2 * L] = 10
    return rec.reshape(-1)
