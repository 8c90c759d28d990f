"""Deterministic random-case generators shared by the parity tests (CPU and GPU).

The shapes follow the fuzz harness described in SURVEY.md Appendix D: reads are pure random,
or carry a mutated copy of the adapter at the start / middle / end (truncated by the read end),
so roughly a third of the cases yield an alignment and every tie-break path is exercised.
"""
import numpy as np

ACGT = "ACGT"
IUPAC = "ACGTRYSWKMBDHVN"


def rand_seq(rng, n, alphabet=ACGT):
    if n <= 0:
        return ""
    idx = rng.integers(0, len(alphabet), size=n)
    return "".join(alphabet[i] for i in idx)


def mutate(rng, seq, sub=0.05, ins=0.02, dele=0.02, alphabet=ACGT):
    out = []
    for ch in seq:
        r = rng.random()
        if r < dele:
            continue
        if r < dele + ins:
            out.append(alphabet[rng.integers(0, len(alphabet))])
        if rng.random() < sub:
            out.append(alphabet[rng.integers(0, len(alphabet))])
        else:
            out.append(ch)
    return "".join(out)


def read_with_adapter(rng, adapter, n, err=None, alphabet=ACGT, n_rate=0.0):
    """One read of (about) length n for `adapter`; mixture of the four shapes."""
    if err is None:
        err = rng.choice([0.0, 0.02, 0.05, 0.1, 0.2])
    shape = rng.integers(0, 5)
    mut = mutate(rng, adapter, sub=err, ins=err / 3, dele=err / 3, alphabet=alphabet) or adapter
    if shape == 0:
        s = rand_seq(rng, n, alphabet)
    elif shape == 1:      # prefix + adapter (possibly truncated by read end)
        p = int(rng.integers(0, max(1, n)))
        s = (rand_seq(rng, p, alphabet) + mut + rand_seq(rng, n, alphabet))[:n]
    elif shape == 2:      # adapter suffix at the read start + tail (5' partial)
        cut = int(rng.integers(0, max(1, len(mut))))
        s = (mut[cut:] + rand_seq(rng, n, alphabet))[:n]
    elif shape == 3:      # adapter prefix exactly at the read end (3' partial)
        keep = int(rng.integers(1, len(mut) + 1))
        s = (rand_seq(rng, max(0, n - keep), alphabet) + mut[:keep])[-n:] if n > 0 else ""
    else:                 # random + adapter + random, full adapter inside
        p = int(rng.integers(0, max(1, n - len(mut) + 1)))
        s = (rand_seq(rng, p, alphabet) + mut + rand_seq(rng, n, alphabet))[:n]
    if n_rate > 0 and s:
        arr = list(s)
        for i in range(len(arr)):
            if rng.random() < n_rate:
                arr[i] = "N"
        s = "".join(arr)
    return s


def locate_cases(seed, count, max_m=64, max_n=230, alphabet=ACGT, adapter_alphabet=None):
    """Yield dicts of Aligner.locate arguments."""
    rng = np.random.default_rng(seed)
    flag_sets = [14, 11, 8, 2, 15, 9, 0, 1, 4, 7, 13, 6, 3, 5, 10, 12]
    for c in range(count):
        m = int(rng.integers(1, max_m + 1)) if rng.random() < 0.2 else int(rng.integers(5, max_m + 1))
        n = int(rng.integers(0, max_n + 1)) if rng.random() < 0.1 else int(rng.integers(10, max_n + 1))
        adapter = rand_seq(rng, m, adapter_alphabet or alphabet)
        flags = flag_sets[int(rng.integers(0, 6))] if rng.random() < 0.9 else flag_sets[int(rng.integers(0, 16))]
        yield dict(
            reference=adapter,
            query=read_with_adapter(rng, adapter, n, alphabet=alphabet),
            max_error_rate=float(rng.choice([0.0, 0.05, 0.1, 0.12, 0.15, 0.2, 0.25, 0.3, 0.5, 1.0])),
            flags=int(flags),
            min_overlap=int(rng.choice([1, 3, 5, 10])),
            indel_cost=int(rng.choice([1, 1, 1, 2, 3, 100000])),
        )


def wildcard_locate_cases(seed, count, max_m=40, max_n=120):
    """Cases for the IUPAC compare modes: wildcards (incl. X) in the adapter, N/R/Y and lower case in the read."""
    rng = np.random.default_rng(seed)
    ad_alpha = "ACGTACGTACGTNRYSWKMBDHVX"
    rd_alpha = "ACGTACGTACGTACGTNRYacgtn"
    for c in range(count):
        m = int(rng.integers(3, max_m + 1))
        n = int(rng.integers(5, max_n + 1))
        adapter = rand_seq(rng, m, ad_alpha)
        # plant with the plain-ACGT projection of the adapter so that matches happen
        proj = "".join(ch if ch in ACGT else ACGT[int(rng.integers(0, 4))] for ch in adapter)
        q = read_with_adapter(rng, proj, n)
        arr = list(q)
        for i in range(len(arr)):
            if rng.random() < 0.05:
                arr[i] = rd_alpha[int(rng.integers(0, len(rd_alpha)))]
        wr, wq = [(True, False), (False, True), (True, True)][int(rng.integers(0, 3))]
        yield dict(
            reference=adapter, query="".join(arr),
            max_error_rate=float(rng.choice([0.0, 0.1, 0.2, 0.3])),
            flags=int(rng.choice([14, 11, 8, 2, 15, 9])),
            wildcard_ref=wr, wildcard_query=wq,
            min_overlap=int(rng.choice([1, 3, 5])),
            indel_cost=int(rng.choice([1, 1, 3, 100000])),
        )


def insert_pairs(seed, count, adapter1, adapter2, lengths=(50, 100, 150), err=0.02, n_rate=0.002,
                 unequal=0.1, lowcomplex=0.05):
    """Paired reads for InsertAligner: fragment shorter than / about / longer than the read length,
    substitutions, Ns, some unequal mate lengths and some low-complexity mates (hit the 100-cap)."""
    rng = np.random.default_rng(seed)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}

    def rc(s):
        return "".join(comp[c] for c in reversed(s))

    def noisy(s):
        arr = list(s)
        for i in range(len(arr)):
            r = rng.random()
            if r < err:
                arr[i] = ACGT[int(rng.integers(0, 4))]
            elif r < err + n_rate:
                arr[i] = "N"
        return "".join(arr)

    for c in range(count):
        L = int(rng.choice(lengths))
        r = rng.random()
        if r < lowcomplex:
            alpha = rng.choice(["A", "AC", "AT", "ACG"])
            frag = rand_seq(rng, int(rng.integers(5, 2 * L)), alpha)
        else:
            flen = int(rng.integers(1, L)) if rng.random() < 0.5 else int(rng.integers(L, 3 * L))
            frag = rand_seq(rng, flen)
        r1 = (frag + adapter1 + rand_seq(rng, L))[:L]
        r2 = (rc(frag) + adapter2 + rand_seq(rng, L))[:L]
        r1, r2 = noisy(r1), noisy(r2)
        if rng.random() < unequal:
            cut = int(rng.integers(1, L))
            if rng.random() < 0.5:
                r1 = r1[:cut]
            else:
                r2 = r2[:cut]
        yield r1, r2


_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "R": "Y", "Y": "R", "a": "t", "c": "g", "g": "c", "t": "a", "n": "n"}


def _rc(seq):
    return "".join(_COMP[c] for c in reversed(seq))


def merge_cases(seed, count):
    """Pairs for MergeOverlapping: both reads sequenced from one fragment (read 1 from its start, read 2 from the
    other strand's start), with substitutions / indels / Ns, so that the reads overlap by anything from nothing to
    everything; some unrelated pairs, very short and empty reads, lower case, IUPAC codes and bytes
    reverse_complement rejects; qualities that decide the error correction either way."""
    rng = np.random.default_rng(seed)
    quals = "#(-2:AFIJ"
    out = []
    for t in range(count):
        L1 = int(rng.choice([0, 1, 2, 5, 12, 30, 50, 75, 100, 150, 250]) if rng.random() < 0.3 else rng.integers(20, 160))
        L2 = L1 if rng.random() < 0.6 else int(rng.integers(0, 200))
        shape = rng.random()
        if shape < 0.12:                                   # unrelated reads
            s1, s2 = rand_seq(rng, L1), rand_seq(rng, L2)
        else:
            lo = max(1, min(L1, L2) // 4)
            F = int(rng.integers(lo, max(lo + 1, L1 + L2 + 30)))
            frag = rand_seq(rng, F, "ACGT" if rng.random() < 0.9 else "AC")      # low complexity: many equally good alignments
            err = float(rng.choice([0.0, 0.0, 0.01, 0.03, 0.08, 0.15]))
            indel = err / 4 if rng.random() < 0.5 else 0.0
            s1 = mutate(rng, frag, sub=err, ins=indel, dele=indel)[:L1]
            s2 = mutate(rng, _rc(frag), sub=err, ins=indel, dele=indel)[:L2]
        def spice(s):
            s = list(s)
            r = rng.random()
            for i in range(len(s)):
                if rng.random() < 0.01:
                    s[i] = "N"
            if r < 0.04 and s:
                for i in range(len(s)):
                    if rng.random() < 0.3:
                        s[i] = s[i].lower()
            elif r < 0.07 and s:
                s[int(rng.integers(0, len(s)))] = str(rng.choice(["R", "Y", "X", "U", ".", "-"]))
            return "".join(s)
        s1, s2 = spice(s1), spice(s2)
        qmode = rng.random()
        def qual(n):
            if qmode < 0.3:
                return "I" * n
            return "".join(quals[i] for i in rng.integers(0, len(quals), size=n))
        out.append(dict(seq1=s1, seq2=s2, qual1=qual(len(s1)), qual2=qual(len(s2)),
                        insert_matched=bool(rng.random() < 0.25),
                        min_overlap=float(rng.choice([0.9, 0.9, 0.5, 0.25, 1.0, 1.5, 2.0, 10.0, 30.0, 0.05])),
                        error_rate=float(rng.choice([0.1, 0.2, 0.2, 0.05, 0.0, 0.3])),
                        mismatch_action=[None, None, "liberal", "conservative", "N"][int(rng.integers(0, 5))]))
    return out


def band_cases(seed, adapter, count):
    """Reads for the banded stage of the fast path: the adapter (mutated, also with indels) somewhere in the read, or a
    prefix of it near the read end, sometimes with a second partial copy in the middle -- the shapes that decide which
    diagonals the band must cover (piece hits on several diagonals, last-column candidates, pieces of an adapter that
    sticks out of the read)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        L = int(rng.integers(20, 170))
        if rng.random() < 0.5:
            err = float(rng.choice([0.0, 0.02, 0.05, 0.1, 0.15, 0.25]))
            read = read_with_adapter(rng, adapter, L, err=err, alphabet="ACGT" if rng.random() < 0.8 else "AC")
        else:
            P = int(rng.integers(1, len(adapter) + 1))
            mut = mutate(rng, adapter[:P], sub=float(rng.choice([0, 0.05, 0.1])), ins=float(rng.choice([0, 0.03, 0.08])),
                         dele=float(rng.choice([0, 0.03, 0.08]))) or adapter[:P]
            pre = rand_seq(rng, max(0, L - len(mut) - int(rng.integers(0, 4))))
            read = (pre + mut + rand_seq(rng, 3))[:L]
            if rng.random() < 0.2:
                read = read[:len(read) // 2] + adapter[:int(rng.integers(5, len(adapter)))] + read[len(read) // 2:]
            read = read[:L]
        out.append(read)
    return out
