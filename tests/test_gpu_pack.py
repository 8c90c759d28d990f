"""The 4-bit packers on the GPU (k_pack through atr_pack_device) against a plain restatement of pack_word
(csrc/adapter_build.hpp): every alignment of a read inside the batch, words that take the A/C/G/T fast path and words
that must not (N, IUPAC, lower case with and without case folding, other bytes, partial last words, the last bytes of
the buffer). Bit-exact codes, lengths and escape flags."""
import numpy as np
import pytest

from atropos_b200 import engine
from atropos_b200.align import _IUPAC_TABLE

pytestmark = pytest.mark.gpu

EXACT = set(b"ACGTRYSWKMBDHVNX")


def _expected(reads, fold_case):
    words, lens = [], []
    for r in reads:
        esc = False
        rw = []
        for w0 in range(0, len(r), 8):
            w = 0
            for t, c in enumerate(r[w0:w0 + 8]):
                if fold_case and 97 <= c <= 122:
                    c -= 32
                esc = esc or c not in EXACT
                w |= _IUPAC_TABLE[c] << (4 * t)
            rw.append(w)
        words.append(rw)
        lens.append(len(r) | (0x8000 if esc else 0))
    return words, lens


@pytest.mark.parametrize("fold_case", [0, 1])
def test_pack_device(fold_case):
    import torch
    rng = np.random.default_rng(11 + fold_case)
    alphabets = [b"ACGT", b"ACGT" * 8 + b"N", b"ACGTNacgtnRYKMX", b"ACGT" * 4 + b"U.-*@"]
    reads = []
    for _ in range(4000):
        alpha = alphabets[int(rng.integers(0, len(alphabets)))]
        n = int(rng.choice([0, 1, 7, 8, 9, 16, 150, 151])) if rng.random() < 0.3 else int(rng.integers(0, 170))
        reads.append(bytes(alpha[i] for i in rng.integers(0, len(alpha), size=n)))
    reads.append(b"ACGTACGT")                       # a full fast-path word as the very last bytes of the buffer
    ascii, offsets = engine.encode_reads(reads)
    n = len(reads)
    ctx = engine.default_context(0)
    dev = torch.device("cuda:0")
    d_ascii = torch.from_numpy(ascii.copy()).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    nwords = int(sum((len(r) + 7) // 8 for r in reads))
    codes = torch.zeros(nwords + 8, dtype=torch.int32, device=dev)
    woff = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    lens = torch.zeros(n, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    engine._lib.check(ctx._L.atr_pack_device(ctx.handle, d_ascii.data_ptr(), d_off.data_ptr(), n, fold_case, codes.data_ptr(),
                                             woff.data_ptr(), lens.data_ptr()), ctx.handle)
    ctx.sync()
    exp_words, exp_lens = _expected(reads, fold_case)
    got_codes = codes.cpu().numpy().view(np.uint32)
    got_woff = woff.cpu().numpy().view(np.uint32)
    got_lens = lens.cpu().numpy().view(np.uint16)
    assert got_woff[0] == 0 and int(got_woff[n]) == nwords
    assert np.array_equal(got_lens, np.array(exp_lens, dtype=np.uint16))
    flat = np.array([w for rw in exp_words for w in rw], dtype=np.uint32)
    assert np.array_equal(np.diff(got_woff.astype(np.int64)), np.array([len(rw) for rw in exp_words]))
    assert np.array_equal(got_codes[:nwords], flat)
