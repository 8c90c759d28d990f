"""atr_pack_reads_host (csrc/host_pack.cpp, no GPU involved): bit-exact codes, word offsets, lengths and escape flags
against a plain restatement of the packing rule, for every alphabet class and alignment, with the AVX2 fast path
(32-base blocks of upper-case A/C/G/T/N) and the scalar path mixed inside one read; threaded and single-threaded."""
import numpy as np
import pytest

from atropos_b200 import engine
from atropos_b200.align import _IUPAC_TABLE

EXACT = set(b"ACGTRYSWKMBDHVNX")


def expected(reads, fold_case):
    words, lens = [], []
    for r in reads:
        esc = False
        for w0 in range(0, len(r), 8):
            w = 0
            for t, c in enumerate(r[w0:w0 + 8]):
                if fold_case and 97 <= c <= 122:
                    c -= 32
                esc = esc or c not in EXACT
                w |= _IUPAC_TABLE[c] << (4 * t)
            words.append(w)
        lens.append(len(r) | (0x8000 if esc else 0))
    return np.array(words, dtype=np.uint32), np.array(lens, dtype=np.uint16)


@pytest.mark.parametrize("fold_case", [0, 1])
@pytest.mark.parametrize("threads", [1, 0])
def test_pack_reads_host(fold_case, threads):
    rng = np.random.default_rng(41 + fold_case)
    alphabets = [b"ACGT", b"ACGT" * 8 + b"N", b"ACGTNacgtnRYKMX", b"ACGT" * 4 + b"U.-*@\x00\x01\x11\x41", bytes(range(256))]
    reads = []
    for _ in range(6000):
        alpha = alphabets[int(rng.integers(0, len(alphabets)))]
        n = int(rng.choice([0, 1, 7, 8, 9, 31, 32, 33, 64, 150, 151])) if rng.random() < 0.3 else int(rng.integers(0, 330))
        reads.append(bytes(alpha[i] for i in rng.integers(0, len(alpha), size=n)))
    ascii, offsets = engine.encode_reads(reads)
    codes, woff, lens = engine.pack_reads_host(ascii, offsets, fold_case=bool(fold_case), threads=threads)
    exp_codes, exp_lens = expected(reads, fold_case)
    assert np.array_equal(lens, exp_lens)
    exp_woff = np.concatenate([[0], np.cumsum([(len(r) + 7) // 8 for r in reads])]).astype(np.uint32)
    assert np.array_equal(woff, exp_woff)
    assert np.array_equal(codes[:exp_codes.size], exp_codes)


def test_pack_reads_host_limits():
    too_long = np.zeros(40000, dtype=np.uint8) + 65
    with pytest.raises(ValueError):
        engine.pack_reads_host(too_long, np.array([0, 40000], dtype=np.int64))
    codes, woff, lens = engine.pack_reads_host(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64))
    assert woff.tolist() == [0] and lens.size == 0
