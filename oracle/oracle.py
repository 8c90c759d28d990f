"""CPU oracle for the adapter-alignment hot path (TEST INFRASTRUCTURE ONLY).

This module restates, on the CPU, the *Python-level* slices of the reference's hot path on top
of the C restatement in ``oracle/atropos_oracle.c``.  The product package ``atropos_b200`` never
imports it; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg
do, and there only as the checker.

Parity is PINNED (not "unpinned"): ``tests/test_oracle_vs_reference.py`` compares every function
here with the real reference (``/root/reference`` imported through ``oracle/ref_loader.py``) on
random inputs, ``tests/test_oracle_golden.py`` checks the reference test-suite's known answers and
the committed golden vectors in ``tests/golden/`` produced by the real reference
(``tests/golden/make_golden.py``).

Restated (citations relative to /root/reference):
  RandomMatchProbability      atropos/util/__init__.py:104-174
  reverse_complement          atropos/util/__init__.py:67-88, 479-482
  compare_suffixes            atropos/align/__init__.py:28-44
  adapter_match_to            atropos/adapters/__init__.py:338-400  (Adapter.match_to; __init__ :259-322)
  linked_match_to             atropos/adapters/__init__.py:671-690  (LinkedAdapter.match_to)
  best_match                  atropos/commands/trim/modifiers.py:107-122 (AdapterCutter._best_match)
  match_insert                atropos/align/__init__.py:250-377     (InsertAligner.match_insert)
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "liboracle.so")
_SRC_PATH = os.path.join(HERE, "atropos_oracle.c")

START_WITHIN_SEQ1, START_WITHIN_SEQ2, STOP_WITHIN_SEQ1, STOP_WITHIN_SEQ2 = 1, 2, 4, 8
SEMIGLOBAL = 15
# adapter types, atropos/adapters/__init__.py:41-56
BACK = START_WITHIN_SEQ2 | STOP_WITHIN_SEQ2 | STOP_WITHIN_SEQ1
FRONT = START_WITHIN_SEQ2 | STOP_WITHIN_SEQ2 | START_WITHIN_SEQ1
PREFIX = STOP_WITHIN_SEQ2
SUFFIX = START_WITHIN_SEQ2
ANYWHERE = SEMIGLOBAL


def build_lib(force=False):
    """Compile oracle/atropos_oracle.c -> oracle/liboracle.so (gcc -O2, same level as the reference build)."""
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC_PATH)):
        return _LIB_PATH
    cc = os.environ.get("CC", "gcc")
    subprocess.check_call([cc, "-O2", "-fPIC", "-shared", "-pthread", "-o", _LIB_PATH, _SRC_PATH])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH) or (
                os.path.exists(_SRC_PATH) and os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC_PATH)):
            build_lib()
        L = ctypes.CDLL(_LIB_PATH)
        c_int_p = ctypes.POINTER(ctypes.c_int)
        L.orc_locate.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_double,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p]
        L.orc_locate.restype = ctypes.c_int
        L.orc_multi_locate.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                       ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p]
        L.orc_multi_locate.restype = ctypes.c_int
        L.orc_compare_prefixes.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, c_int_p]
        L.orc_compare_prefixes.restype = None
        L.orc_locate_batch.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        L.orc_locate_batch.restype = ctypes.c_int
        L.orc_init()
        _lib = L
    return _lib


def _b(s):
    return s if isinstance(s, bytes) else s.encode("ascii")


# --------------------------------------------------------------------------------------------
# native-level functions (C restatement)
# --------------------------------------------------------------------------------------------

def locate(reference, query, max_error_rate, flags=SEMIGLOBAL, wildcard_ref=False, wildcard_query=False,
           min_overlap=1, indel_cost=1):
    """Aligner(reference, ...).locate(query) -> None | 6-tuple   (_align.pyx:266-491)"""
    r, q = _b(reference), _b(query)
    out = (ctypes.c_int * 6)()
    found = lib().orc_locate(r, len(r), q, len(q), float(max_error_rate), int(flags), int(bool(wildcard_ref)),
                             int(bool(wildcard_query)), int(min_overlap), int(indel_cost), out)
    if found < 0:
        raise MemoryError()
    return tuple(out) if found else None


def multi_locate(reference, query, max_error_rate, flags=SEMIGLOBAL, min_overlap=1, max_matches=100):
    """MultiAligner(rate, flags, min_overlap).locate(reference, query, max_matches) (_align.pyx:593-772)"""
    r, q = _b(reference), _b(query)
    out = (ctypes.c_int * (6 * (max_matches + 2 + len(r))))()
    cnt = lib().orc_multi_locate(r, len(r), q, len(q), float(max_error_rate), int(flags), int(min_overlap),
                                 int(max_matches), out)
    if cnt < 0:
        raise MemoryError()
    if cnt == 0:
        return None
    return [tuple(out[6 * t:6 * t + 6]) for t in range(cnt)]


def compare_prefixes(ref, query, wildcard_ref=False, wildcard_query=False):
    """_align.pyx:501-544"""
    r, q = _b(ref), _b(query)
    out = (ctypes.c_int * 6)()
    lib().orc_compare_prefixes(r, len(r), q, len(q), int(bool(wildcard_ref)), int(bool(wildcard_query)), out)
    return tuple(out)


def compare_suffixes(suffix_ref, suffix_query, wildcard_ref=False, wildcard_query=False):
    """align/__init__.py:28-44"""
    sr, sq = suffix_ref[::-1], suffix_query[::-1]
    _, length, _, _, matches, errors = compare_prefixes(sr, sq, wildcard_ref, wildcard_query)
    return (len(sr) - length, len(sr), len(sq) - length, len(sq), matches, errors)


def locate_batch(reference, reads_concat, offsets, max_error_rate, flags, wildcard_ref=False,
                 wildcard_query=False, min_overlap=1, indel_cost=1, threads=1):
    """Vector form: reads_concat = uint8 array of all reads back to back, offsets = int64[n+1].
    Returns int32 array [n, 7] = (found, refstart, refstop, querystart, querystop, matches, errors)."""
    reads_concat = np.ascontiguousarray(reads_concat, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    out = np.zeros((n, 7), dtype=np.int32)
    r = _b(reference)
    rc = lib().orc_locate_batch(r, len(r), float(max_error_rate), int(flags), int(bool(wildcard_ref)),
                                int(bool(wildcard_query)), int(min_overlap), int(indel_cost),
                                reads_concat.ctypes.data, offsets.ctypes.data, n, int(threads), out.ctypes.data)
    if rc != 0:
        raise RuntimeError("orc_locate_batch failed")
    return out


# --------------------------------------------------------------------------------------------
# util restatements
# --------------------------------------------------------------------------------------------

class RandomMatchProbability(object):
    """Binomial-tail random-match probability with the reference's exact arithmetic
    (util/__init__.py:104-174): big-int factorials, true division first with an
    ``OverflowError`` fallback to floor division, terms accumulated from i = matches upward."""

    def __init__(self):
        self.cache = {}
        self.fact = [1, 1]

    def factorial(self, num):
        f = self.fact
        while len(f) <= num:
            f.append(f[-1] * len(f))
        return f[num]

    def __call__(self, matches, size, match_prob=0.25, mismatch_prob=0.75):
        key = (matches, size, match_prob)            # the reference's key ignores mismatch_prob (:129)
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


def _complement_table():
    """util/__init__.py:67-88"""
    nuc = {'A': 'T', 'C': 'G', 'R': 'Y', 'S': 'S', 'W': 'W', 'K': 'M', 'B': 'V', 'D': 'H', 'N': 'N'}
    for base, comp in tuple(nuc.items()):
        nuc[comp] = base
        nuc[base.lower()] = comp.lower()
        nuc[comp.lower()] = base.lower()
    return nuc


_COMPLEMENT = _complement_table()
IUPAC_BASES = frozenset(('X',) + tuple(_COMPLEMENT.keys()))


def reverse_complement(seq):
    """util/__init__.py:479-482 -- raises KeyError on any byte outside the IUPAC table."""
    return "".join(_COMPLEMENT[b] for b in reversed(seq))


# --------------------------------------------------------------------------------------------
# Adapter.match_to and friends
# --------------------------------------------------------------------------------------------

class OracleAdapter(object):
    """The state `Adapter.__init__` derives (adapters/__init__.py:259-322), minus statistics."""

    def __init__(self, sequence, where, max_error_rate=0.1, min_overlap=3, read_wildcards=False,
                 adapter_wildcards=True, indels=True, indel_cost=1, match_probability=None, max_rmp=None):
        if len(sequence) == 0:
            raise ValueError("Empty adapter sequence")
        sequence = sequence.upper().replace('U', 'T')     # parse_braces is out of scope (spec parsing)
        seq_set = set(sequence)
        if seq_set <= set('ACGT'):
            adapter_wildcards = False
        if adapter_wildcards and not seq_set <= IUPAC_BASES:
            raise ValueError("Invalid character(s) in adapter sequence")
        self.sequence = sequence
        self.where = where
        self.max_error_rate = max_error_rate
        self.min_overlap = min(min_overlap, len(sequence))
        self.match_probability = match_probability
        self.max_rmp = max_rmp
        self.indels = indels
        self.adapter_wildcards = adapter_wildcards
        self.read_wildcards = read_wildcards
        self.front_flag = None if where == ANYWHERE else (where not in (BACK, SUFFIX))
        self.indel_cost = indel_cost if indels else 100000

    def match_to(self, read_sequence):
        """Adapter.match_to (adapters/__init__.py:338-400).
        Returns None or (astart, astop, rstart, rstop, matches, errors, front)."""
        read_seq = read_sequence.upper()
        pos = -1
        if not self.adapter_wildcards:
            if self.where == PREFIX:
                if read_seq.startswith(self.sequence):
                    pos = 0
            elif self.where == SUFFIX:
                if read_seq.endswith(self.sequence):
                    pos = len(read_seq) - len(self.sequence)
            else:
                pos = read_seq.find(self.sequence)
        if pos >= 0:
            m = len(self.sequence)
            return self._make(0, m, pos, pos + m, m, 0)
        if not self.indels and self.where in (PREFIX, SUFFIX):
            if self.where == PREFIX:
                alignment = compare_prefixes(self.sequence, read_seq, self.adapter_wildcards, self.read_wildcards)
            else:
                alignment = compare_suffixes(self.sequence, read_seq, self.adapter_wildcards, self.read_wildcards)
        else:
            alignment = locate(self.sequence, read_seq, self.max_error_rate, self.where, self.adapter_wildcards,
                               self.read_wildcards, self.min_overlap, self.indel_cost)
        if alignment:
            astart, astop, rstart, rstop, matches, errors = alignment
            size = astop - astart
            if ((size >= self.min_overlap and errors / size <= self.max_error_rate) and
                    (self.max_rmp is None or self.match_probability(matches, size) <= self.max_rmp)):
                return self._make(astart, astop, rstart, rstop, matches, errors)
        return None

    def _make(self, astart, astop, rstart, rstop, matches, errors):
        # Match.__init__ (align/__init__.py:70-88)
        front = (rstart == 0) if self.front_flag is None else self.front_flag
        length = astop - astart
        if length <= 0:
            raise ValueError('Match length must be >= 0')
        if length - errors <= 0:
            raise ValueError('A Match requires at least one matching position.')
        return (astart, astop, rstart, rstop, matches, errors, front)


def best_match(adapters, read_sequence):
    """AdapterCutter._best_match (modifiers.py:107-122): strictly more matches wins, first adapter on ties.
    Returns (adapter_index, match) or None."""
    best = None
    for idx, adapter in enumerate(adapters):
        match = adapter.match_to(read_sequence)
        if match is None:
            continue
        if best is None or match[4] > best[1][4]:
            best = (idx, match)
    return best


def linked_match_to(front_adapter, back_adapter, read_sequence):
    """LinkedAdapter.match_to (adapters/__init__.py:671-690).
    Returns None or (front_match, back_match|None); back coordinates are relative to read[front.rstop:]."""
    fm = front_adapter.match_to(read_sequence)
    if fm is None:
        return None
    bm = back_adapter.match_to(read_sequence[fm[3]:])
    return (fm, bm)


# --------------------------------------------------------------------------------------------
# InsertAligner.match_insert
# --------------------------------------------------------------------------------------------

class OracleInsertAligner(object):
    """InsertAligner (align/__init__.py:178-377)."""

    def __init__(self, adapter1, adapter2, match_probability=None, insert_max_rmp=1E-6, adapter_max_rmp=0.001,
                 min_insert_overlap=1, max_insert_mismatch_frac=0.2, min_adapter_overlap=1,
                 max_adapter_mismatch_frac=0.2, adapter_check_cutoff=9, base_probs=None,
                 adapter_wildcards=True, read_wildcards=False):
        self.adapter1, self.adapter2 = adapter1, adapter2
        self.match_probability = match_probability or RandomMatchProbability()
        self.insert_max_rmp = insert_max_rmp
        self.adapter_max_rmp = adapter_max_rmp
        self.min_insert_overlap = min_insert_overlap
        self.max_insert_mismatch_frac = float(max_insert_mismatch_frac)
        self.min_adapter_overlap = min_adapter_overlap
        self.max_adapter_mismatch_frac = float(max_adapter_mismatch_frac)
        self.adapter_check_cutoff = adapter_check_cutoff
        self.base_probs = base_probs or dict(match_prob=0.25, mismatch_prob=0.75)
        self.adapter_wildcards = adapter_wildcards
        self.read_wildcards = read_wildcards
        # NB :230-233 hands the *unconverted* max_insert_mismatch_frac to MultiAligner (a C double either way)
        self._rate = float(max_insert_mismatch_frac)

    def match_insert(self, seq1, seq2):
        """Returns None or (insert_match 6-tuple, m1, m2) with m = None | (astart, astop, rstart, rstop, matches, errors)."""
        seq_len1, seq_len2 = len(seq1), len(seq2)
        seq_len = min(seq_len1, seq_len2)
        if seq_len1 > seq_len2:
            seq1 = seq1[:seq_len2]
        elif seq_len2 > seq_len1:
            seq2 = seq2[:seq_len1]
        seq2_rc = reverse_complement(seq2)

        def _match(insert_match, offset, insert_match_size, _prob):
            if offset < self.min_adapter_overlap:
                return (insert_match, None, None)

            def _adapter_match(insert_seq, adapter_seq):
                amatch = compare_prefixes(insert_seq[insert_match_size:], adapter_seq,
                                          wildcard_ref=self.adapter_wildcards, wildcard_query=self.read_wildcards)
                alen = min(offset, len(adapter_seq))
                return amatch, alen, round(alen * self.max_adapter_mismatch_frac)

            a1_match, a1_length, a1_max = _adapter_match(seq1, self.adapter1)
            a2_match, a2_length, a2_max = _adapter_match(seq2, self.adapter2)
            if a1_match[5] > a1_max and a2_match[5] > a2_max:
                return None
            if min(a1_length, a2_length) > self.adapter_check_cutoff:
                a1_prob = self.match_probability(a1_match[4], a1_length)
                a2_prob = self.match_probability(a2_match[4], a2_length)
                if (a1_prob * a2_prob) > self.adapter_max_rmp:
                    return None
            mismatches = min(a1_match[5], a2_match[5])

            def _create(alen, slen):
                alen = min(alen, slen - insert_match_size)
                mm = min(alen, mismatches)
                if alen <= 0 or alen - mm <= 0:                     # Match.__init__ :85-88
                    raise ValueError('invalid Match')
                return (0, alen, insert_match_size, slen, alen - mm, mm)

            return (insert_match, _create(a1_length, seq_len1), _create(a2_length, seq_len2))

        insert_matches = multi_locate(seq2_rc, seq1, self._rate, START_WITHIN_SEQ1 | STOP_WITHIN_SEQ2,
                                      self.min_insert_overlap)
        if insert_matches:
            filtered = []
            for im in insert_matches:
                offset = min(im[0], seq_len - im[3])
                size = seq_len - offset
                prob = self.match_probability(im[4], size, **self.base_probs)
                if prob <= self.insert_max_rmp:
                    filtered.append((im, offset, size, prob))
            if filtered:
                if len(filtered) == 1:
                    return _match(*filtered[0])
                filtered.sort(key=lambda x: x[3])
                for args in filtered:
                    match = _match(*args)
                    if match:
                        return match
        return None
