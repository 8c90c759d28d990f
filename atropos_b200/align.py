"""GPU-backed mirror of the reference's `atropos.align` / `atropos.align._align` interface.

Same names, argument meaning and error behaviour as the reference (citations relative to the
reference checkout):

    Aligner, MultiAligner, compare_prefixes, locate     atropos/align/_align.pyx
    compare_suffixes, Match, MatchInfo, InsertAligner   atropos/align/__init__.py
    START_WITHIN_SEQ1 ... SEMIGLOBAL                    atropos/align/__init__.py:17-26

Every call goes through the C ABI (`libatropos_b200.so`) to the CUDA kernels; there is no CPU
implementation in this package. The per-call methods (`locate(query)`, `match_insert(seq1, seq2)`)
are batches of one and exist for drop-in compatibility; the batched twins (`locate_batch`,
`match_insert_batch`) are what a pipeline should call.
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from . import _abi, engine
from .util import RandomMatchProbability, reverse_complement  # noqa: F401  (re-exported like the reference)

START_WITHIN_SEQ1 = 1
START_WITHIN_SEQ2 = 2
STOP_WITHIN_SEQ1 = 4
STOP_WITHIN_SEQ2 = 8
SEMIGLOBAL = START_WITHIN_SEQ1 | START_WITHIN_SEQ2 | STOP_WITHIN_SEQ1 | STOP_WITHIN_SEQ2


def _tuple_of(rec):
    return (int(rec["astart"]), int(rec["astop"]), int(rec["rstart"]), int(rec["rstop"]), int(rec["matches"]),
            int(rec["errors"]))


class Aligner(object):
    """Aligner(reference, max_error_rate, flags=SEMIGLOBAL, wildcard_ref=False, wildcard_query=False,
    min_overlap=1, indel_cost=1)  --  _align.pyx:121-494."""

    def __init__(self, reference, max_error_rate, flags=SEMIGLOBAL, wildcard_ref=False, wildcard_query=False,
                 min_overlap=1, indel_cost=1, device=0):
        self.max_error_rate = float(max_error_rate)
        self.flags = int(flags)
        self.wildcard_ref = bool(wildcard_ref)
        self.wildcard_query = bool(wildcard_query)
        self._device = device
        self._set = None
        self._dpmatrix = None
        self.debug = False
        self.reference = reference
        self.min_overlap = min_overlap
        self.indel_cost = indel_cost

    def __reduce__(self):
        return (Aligner, (self.str_reference, self.max_error_rate, self.flags, self.wildcard_ref, self.wildcard_query,
                          self._min_overlap, self._indel_cost))

    @property
    def min_overlap(self):
        return self._min_overlap

    @min_overlap.setter
    def min_overlap(self, value):
        value = int(value)
        if value < 1:
            raise ValueError('Minimum overlap must be at least 1')
        self._min_overlap = value
        self._set = None

    def _set_indel_cost(self, value):
        if value < 1:
            raise ValueError('Insertion/deletion cost must be at least 1')
        self._indel_cost = int(value)
        self._set = None

    indel_cost = property(None, _set_indel_cost)      # write-only, like the reference (:223-232)

    @property
    def reference(self):
        """The *translated* reference bytes, like the reference's getter (:234-249)."""
        raw = self.str_reference.encode('ascii')
        if self.wildcard_ref:
            return raw.translate(_IUPAC_TABLE)
        if self.wildcard_query:
            return raw.translate(_ACGT_TABLE)
        return raw

    @reference.setter
    def reference(self, reference):
        reference.encode('ascii')
        self.str_reference = reference
        self.m = len(reference)
        self._set = None

    @property
    def dpmatrix(self):
        return self._dpmatrix

    def enable_debug(self):
        """The DP matrix is never materialised on the GPU; debugging output is not available."""
        self.debug = True

    def _adapterset(self):
        key = engine.context_key(self._device)         # a set belongs to one context = one (process, thread)
        if self._set is None or self._set_key != key:
            ctx = engine.default_context(self._device)
            desc = _abi.make_adapter_desc(self.str_reference, self.max_error_rate, self.flags, self.wildcard_ref,
                                          self.wildcard_query, self._min_overlap, self._indel_cost)
            self._set = engine.AdapterSet(ctx, [desc])
            self._set_key = key
        return self._set

    def locate(self, query):
        """locate(query) -> None | (refstart, refstop, querystart, querystop, matches, errors)   (:266-491)"""
        if self.m == 0:
            # the reference walks an empty matrix and reports no alignment
            query.encode('ascii')
            return None
        res = self.locate_batch([query])
        rec = res[0]
        if rec["status"] == _abi.ATR_ST_NONE:
            return None
        tup = _tuple_of(rec)
        assert tup[1] - tup[0] > 0
        return tup

    def locate_batch(self, reads, win=None):
        """Batched twin: reads = list of str, or (ascii uint8 array, int64 offsets). Returns a MATCH_DTYPE array."""
        ascii, offsets = reads if isinstance(reads, tuple) else engine.encode_reads(reads)
        return self._adapterset().locate_host(ascii, offsets, win=win, fold_case=False)


def locate(reference, query, max_error_rate, flags=SEMIGLOBAL, wildcard_ref=False, wildcard_query=False,
           min_overlap=1):
    """_align.pyx:496-499"""
    aligner = Aligner(reference, max_error_rate, flags, wildcard_ref, wildcard_query)
    aligner.min_overlap = min_overlap
    return aligner.locate(query)


def compare_prefixes(ref, query, wildcard_ref=False, wildcard_query=False):
    """_align.pyx:501-544"""
    return engine.default_context().compare_prefixes(ref, query, wildcard_ref, wildcard_query)


def compare_suffixes(suffix_ref, suffix_query, wildcard_ref=False, wildcard_query=False):
    """align/__init__.py:28-44"""
    suffix_ref = suffix_ref[::-1]
    suffix_query = suffix_query[::-1]
    _, length, _, _, matches, errors = compare_prefixes(suffix_ref, suffix_query, wildcard_ref, wildcard_query)
    return (len(suffix_ref) - length, len(suffix_ref), len(suffix_query) - length, len(suffix_query), matches, errors)


class MultiAligner(object):
    """MultiAligner(max_error_rate, flags=SEMIGLOBAL, min_overlap=1)  --  _align.pyx:548-787."""

    def __init__(self, max_error_rate, flags=SEMIGLOBAL, min_overlap=1, device=0):
        self.max_error_rate = float(max_error_rate)
        self.flags = int(flags)
        self._min_overlap = int(min_overlap)
        self._device = device

    def __reduce__(self):
        return (MultiAligner, (self.max_error_rate, self.flags, self._min_overlap))

    def locate(self, reference, query, max_matches=100):
        """-> None | list of (refstart, refstop, querystart, querystop, matches, errors)   (:593-772)"""
        return engine.default_context(self._device).multi_locate(reference, query, self.max_error_rate, self.flags,
                                                                 self._min_overlap, max_matches)


class Match(object):
    """An alignment match -- align/__init__.py:51-170 (same slots, same checks)."""
    __slots__ = ['astart', 'astop', 'rstart', 'rstop', 'matches', 'errors', 'front', 'adapter', 'read', 'length']

    def __init__(self, astart, astop, rstart, rstop, matches, errors, front=None, adapter=None, read=None):
        self.astart = astart
        self.astop = astop
        self.rstart = rstart
        self.rstop = rstop
        self.matches = matches
        self.errors = errors
        self.front = self._guess_is_front() if front is None else front
        self.adapter = adapter
        self.read = read
        self.length = self.astop - self.astart
        if self.length <= 0:
            raise ValueError('Match length must be >= 0')
        if self.length - self.errors <= 0:
            raise ValueError('A Match requires at least one matching position.')

    def __repr__(self):
        return ('Match(astart={0}, astop={1}, rstart={2}, rstop={3}, matches={4}, errors={5})').format(
            self.astart, self.astop, self.rstart, self.rstop, self.matches, self.errors)

    def __eq__(self, other):
        return isinstance(other, Match) and self.fields() == other.fields() and self.front == other.front

    def fields(self):
        return (self.astart, self.astop, self.rstart, self.rstop, self.matches, self.errors)

    def copy(self):
        return Match(self.astart, self.astop, self.rstart, self.rstop, self.matches, self.errors, self.front,
                     self.adapter, self.read)

    def _guess_is_front(self):
        return self.rstart == 0

    def wildcards(self, wildcard_char='N'):
        wildcards = [self.read.sequence[self.rstart + i] for i in range(self.length)
                     if (self.adapter.sequence[self.astart + i] == wildcard_char and
                         self.rstart + i < len(self.read.sequence))]
        return ''.join(wildcards)

    def rest(self):
        if self.front:
            return self.read.sequence[:self.rstart]
        return self.read.sequence[self.rstop:]

    def get_info_record(self):
        seq = self.read.sequence
        qualities = self.read.qualities
        if qualities is None:
            qualities = ''
        rsize = rsize_total = self.rstop - self.rstart
        if self.front and self.rstart > 0:
            rsize_total = self.rstop
        elif not self.front and self.rstop < len(seq):
            rsize_total = len(seq) - self.rstart
        return MatchInfo(self.read.name, self.errors, self.rstart, self.rstop, seq[0:self.rstart],
                         seq[self.rstart:self.rstop], seq[self.rstop:], self.adapter.name, qualities[0:self.rstart],
                         qualities[self.rstart:self.rstop], qualities[self.rstop:], self.front,
                         self.astop - self.astart, rsize, rsize_total)


MatchInfo = namedtuple("MatchInfo", (
    "read_name", "errors", "rstart", "rstop", "seq_before", "seq_adapter", "seq_after", "adapter_name", "qual_before",
    "qual_adapter", "qual_after", "is_front", "asize", "rsize_adapter", "rsize_total"))


def _match_from_record(rec):
    """atr_match record -> Match | None; raises what Match.__init__ would."""
    st = int(rec["status"])
    if st == _abi.ATR_ST_NONE:
        return None
    if st == _abi.ATR_ST_INVALID:
        raise ValueError('A Match requires at least one matching position.')
    return Match(*_tuple_of(rec))


class InsertAligner(object):
    """InsertAligner(adapter1, adapter2, ...)  --  align/__init__.py:178-377 (same keyword arguments)."""

    def __init__(self, adapter1, adapter2, match_probability=None, insert_max_rmp=1E-6, adapter_max_rmp=0.001,
                 min_insert_overlap=1, max_insert_mismatch_frac=0.2, min_adapter_overlap=1,
                 max_adapter_mismatch_frac=0.2, adapter_check_cutoff=9, base_probs=None, adapter_wildcards=True,
                 read_wildcards=False, device=0):
        self.adapter1 = adapter1
        self.adapter1_len = len(adapter1)
        self.adapter2 = adapter2
        self.adapter2_len = len(adapter2)
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
        self.aligner = MultiAligner(max_insert_mismatch_frac, START_WITHIN_SEQ1 | STOP_WITHIN_SEQ2, min_insert_overlap,
                                    device=device)
        self._device = device
        self._set = None

    # -- tables ------------------------------------------------------------------------------
    def descriptor(self, max_len):
        """(AtrInsertDesc, keepalive) with the probability tables for reads up to max_len."""
        L = max(int(max_len), 1)
        rate = self.max_insert_mismatch_frac
        kmax = int(rate * L)
        mp = self.match_probability
        ins = np.ones((L + 1, kmax + 1), dtype=np.float64)       # 1.0 > any max_rmp for never-read entries
        for size in range(1, L + 1):
            for cost in range(0, min(int(rate * size), size) + 1):
                ins[size, cost] = mp(size - cost, size, **self.base_probs)
        amax = max(self.adapter1_len, self.adapter2_len)
        adp = np.zeros((amax + 1, amax + 1), dtype=np.float64)
        for alen in range(0, amax + 1):
            for matches in range(0, alen + 1):
                adp[alen, matches] = mp(matches, alen)
        a1, a2 = self.adapter1.encode('ascii'), self.adapter2.encode('ascii')
        d = _abi.AtrInsertDesc()
        d.adapter1, d.adapter1_len, d.adapter2, d.adapter2_len = a1, len(a1), a2, len(a2)
        d.insert_max_rmp, d.adapter_max_rmp = float(self.insert_max_rmp), float(self.adapter_max_rmp)
        d.min_insert_overlap = int(self.min_insert_overlap)
        d.max_insert_mismatch_frac = rate
        d.min_adapter_overlap = int(self.min_adapter_overlap)
        d.max_adapter_mismatch_frac = self.max_adapter_mismatch_frac
        d.adapter_check_cutoff = int(self.adapter_check_cutoff)
        d.adapter_wildcards, d.read_wildcards = int(bool(self.adapter_wildcards)), int(bool(self.read_wildcards))
        d.max_len = L
        d.insert_prob, d.adapter_prob = ins.ctypes.data, adp.ctypes.data
        return d, [a1, a2, ins, adp]

    def _insertset(self, max_len):
        key = engine.context_key(self._device)
        if self._set is not None and self._set_key != key:
            self._set = None
        if self._set is not None and self._set.max_len >= max_len:
            return self._set
        L = max(int(max_len), 1)
        if self._set is not None:
            L = max(L, 2 * self._set.max_len)          # grow geometrically
        d, keep = self.descriptor(L)
        self._set = engine.InsertSet(engine.default_context(self._device), d, keep)
        self._set_key = key
        return self._set

    # -- the reference's per-pair call --------------------------------------------------------------
    def match_insert(self, seq1, seq2):
        """-> None | (insert_match, Match|None, Match|None)   (align/__init__.py:250-377)"""
        res = self.match_insert_batch([seq1], [seq2])
        return self.result_from_record(res[0])

    def match_insert_batch(self, reads1, reads2):
        """Batched twin. reads{1,2}: list of str or (ascii, offsets). Returns an INSERT_DTYPE array."""
        a1, o1 = reads1 if isinstance(reads1, tuple) else engine.encode_reads(reads1)
        a2, o2 = reads2 if isinstance(reads2, tuple) else engine.encode_reads(reads2)
        n = len(o1) - 1
        if n == 0:
            return np.empty(0, dtype=_abi.INSERT_DTYPE)
        max_len = int(min(np.diff(o1).max(), np.diff(o2).max()))
        return self._insertset(max_len).match_insert_host(a1, o1, a2, o2)

    @staticmethod
    def result_from_record(rec):
        st = int(rec["insert"]["status"])
        if st == _abi.ATR_ST_NONE:
            return None
        if st == _abi.ATR_ST_KEYERROR:
            raise KeyError('reverse_complement: base outside the IUPAC alphabet')
        if st != _abi.ATR_ST_MATCH:
            raise RuntimeError("match_insert: unexpected status %d" % st)
        return (_tuple_of(rec["insert"]), _match_from_record(rec["match1"]), _match_from_record(rec["match2"]))


def _tables():
    d = dict(A=1, C=2, G=4, T=8, U=8)
    acgt = bytearray(256)
    for c, v in d.items():
        acgt[ord(c)] = v
        acgt[ord(c.lower())] = v
    A, C_, G, T = 1, 2, 4, 8
    d = dict(X=0, A=A, C=C_, G=G, T=T, U=T, R=A | G, Y=C_ | T, S=G | C_, W=A | T, K=G | T, M=A | C_, B=C_ | G | T,
             D=A | G | T, H=A | C_ | T, V=A | C_ | G, N=A | C_ | G | T)
    iupac = bytearray(256)
    for c, v in d.items():
        iupac[ord(c)] = v
        iupac[ord(c.lower())] = v
    return bytes(acgt), bytes(iupac)


_ACGT_TABLE, _IUPAC_TABLE = _tables()
