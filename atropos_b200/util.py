"""Host-side tables the hot path needs (the reference keeps the same quantities in atropos/util/__init__.py).

Random-match probabilities decide accept / reject in `Adapter.match_to` (max_rmp) and in `InsertAligner`
(insert_max_rmp, adapter_max_rmp); a decision must not flip in the last bit, so every probability is evaluated on
the HOST with the very floating-point expression of the reference (util/__init__.py:117-155) -- Python big-int
factorials, true division with the OverflowError fallback to floor division, the sum accumulated upwards from
`matches` -- and shipped to the GPU as a table. The kernels never redo float math.

Unlike the reference (one dict entry per query, factorials grown on demand) the work is organised by ROW: all
terms of a sequence size are computed once, and a probability is a partial sum over that row.
"""
import math


def binomial_terms(size, match_prob, mismatch_prob):
    """t[i] = mismatch_prob**(size-i) * match_prob**i * C(size, i), each in the reference's order of operations."""
    nfac = math.factorial(size)
    terms = []
    for i in range(size + 1):
        j = size - i
        try:
            div = nfac / math.factorial(i) / math.factorial(j)
        except OverflowError:                      # int too large for a float: the reference falls back to floor division
            div = nfac // math.factorial(i) // math.factorial(j)
        terms.append((mismatch_prob ** j) * (match_prob ** i) * div)
    return terms


class RandomMatchProbability(object):
    """P(at least `matches` of `size` random bases match): callable like the reference's class of the same name,
    `rmp(matches, size, match_prob=0.25, mismatch_prob=0.75)`, with identical results bit for bit."""

    def __init__(self, init_size=150):
        self._rows = {}                            # (size, match_prob) -> terms; like the reference's cache key, the
        #                                            mismatch probability of the FIRST query of a row sticks

    def row(self, size, match_prob=0.25, mismatch_prob=0.75):
        key = (size, match_prob)
        terms = self._rows.get(key)
        if terms is None:
            terms = self._rows[key] = binomial_terms(size, match_prob, mismatch_prob)
        return terms

    def __call__(self, matches, size, match_prob=0.25, mismatch_prob=0.75):
        if matches == size:
            return match_prob ** matches           # util/__init__.py:137-138
        prob = 0.0
        for term in self.row(size, match_prob, mismatch_prob)[matches:]:
            prob += term
        return prob

    def table(self, size, match_prob=0.25, mismatch_prob=0.75):
        """[P(>= m of size) for m = 0..size]"""
        return [self(m, size, match_prob, mismatch_prob) for m in range(size + 1)]


# complement of every IUPAC letter, upper and lower case (util/__init__.py:67-88)
BASE_COMPLEMENTS = {
    'A': 'T', 'T': 'A', 'C': 'G', 'G': 'C', 'R': 'Y', 'Y': 'R', 'S': 'S', 'W': 'W', 'K': 'M', 'M': 'K',
    'B': 'V', 'V': 'B', 'D': 'H', 'H': 'D', 'N': 'N',
    'a': 't', 't': 'a', 'c': 'g', 'g': 'c', 'r': 'y', 'y': 'r', 's': 's', 'w': 'w', 'k': 'm', 'm': 'k',
    'b': 'v', 'v': 'b', 'd': 'h', 'h': 'd', 'n': 'n',
}
IUPAC_BASES = frozenset('X' + ''.join(BASE_COMPLEMENTS))


def reverse_complement(seq):
    """util/__init__.py:479-482. (On the GPU path this is one bit reversal per 8 bases; this host version serves
    callers outside the batch path and raises the same KeyError on a byte outside the table.)"""
    return "".join([BASE_COMPLEMENTS[base] for base in reversed(seq)])


def expand_braces(sequence):
    """'TGA{5}CT' -> 'TGAAAAACT': the repeat notation of adapter sequences (adapters/__init__.py:933-970; same
    ValueErrors). A brace group repeats the ONE character in front of it."""
    out = []
    i, n = 0, len(sequence)
    while i < n:
        ch = sequence[i]
        if ch == '}':
            raise ValueError('"}" cannot be used here' if not out or i == 0 else 'Expected "{"')
        if ch == '{':
            if not out or (i > 0 and sequence[i - 1] == '}'):
                raise ValueError('"{" must be used after a character')
            close = sequence.find('}', i + 1)
            nxt = sequence.find('{', i + 1)
            if close < 0:
                raise ValueError("Unterminated expression")
            if 0 <= nxt < close:
                raise ValueError('"}" expected')
            count = int(sequence[i + 1:close])          # ValueError on a non-number, like int() in the reference
            if not 0 <= count <= 10000:
                raise ValueError('Value {} invalid'.format(count))
            last = out.pop()
            out.append(last[:-1] + last[-1] * count)
            i = close + 1
            continue
        j = i
        while j < n and sequence[j] not in '{}':
            j += 1
        out.append(sequence[i:j])
        i = j
    return ''.join(out)
