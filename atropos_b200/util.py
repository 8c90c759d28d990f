"""Host-side helpers the hot path needs, mirroring atropos/util/__init__.py.

`RandomMatchProbability` is evaluated on the HOST with the reference's exact arithmetic (Python
big-int factorials, true division with an OverflowError fallback to floor division, ascending
accumulation; util/__init__.py:104-174) and shipped to the GPU as tables -- the kernels never
redo float math.
"""


class RandomMatchProbability(object):
    """Same call signature and caching behaviour as the reference class (util/__init__.py:104-174)."""

    def __init__(self, init_size=150):
        self.cache = {}
        self.factorials = [1] * init_size
        self.max_n = 1
        self.cur_array_size = init_size

    def __call__(self, matches, size, match_prob=0.25, mismatch_prob=0.75):
        key = (matches, size, match_prob)
        prob = self.cache.get(key, None)
        if prob:
            return prob
        if matches == size:
            prob = match_prob ** matches
        else:
            nfac = self.factorial(size)
            prob = 0.0
            for i in range(matches, size + 1):
                j = size - i
                try:
                    div = nfac / self.factorial(i) / self.factorial(j)
                except OverflowError:
                    div = nfac // self.factorial(i) // self.factorial(j)
                prob += (mismatch_prob ** j) * (match_prob ** i) * div
        self.cache[key] = prob
        return prob

    def factorial(self, num):
        if num > self.max_n:
            self._fill_upto(num)
        return self.factorials[num]

    def _fill_upto(self, num):
        if num >= self.cur_array_size:
            self.factorials += [1] * (num - self.cur_array_size + 1)
            self.cur_array_size = len(self.factorials)
        idx = self.max_n
        while idx < num:
            self.factorials[idx + 1] = (idx + 1) * self.factorials[idx]
            idx += 1
        self.max_n = idx


def _build_complements():
    nuc = {'A': 'T', 'C': 'G', 'R': 'Y', 'S': 'S', 'W': 'W', 'K': 'M', 'B': 'V', 'D': 'H', 'N': 'N'}
    for base, comp in tuple(nuc.items()):
        nuc[comp] = base
        nuc[base.lower()] = comp.lower()
        nuc[comp.lower()] = base.lower()
    return nuc


BASE_COMPLEMENTS = _build_complements()
IUPAC_BASES = frozenset(('X',) + tuple(BASE_COMPLEMENTS.keys()))


def reverse_complement(seq):
    """util/__init__.py:479-482. (On the GPU path this is a one-instruction bit reversal per 8 bases; this
    host version exists for callers outside the batch path.)"""
    return "".join(BASE_COMPLEMENTS[base] for base in reversed(seq))
