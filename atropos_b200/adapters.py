"""GPU-backed mirror of the alignment-facing part of `atropos.adapters` (reference:
atropos/adapters/__init__.py): adapter type constants (:41-56), `Adapter` (:231-505: constructor
normalisation, `match_to`), `LinkedAdapter.match_to` (:637-690). Spec parsing, the adapter cache,
colorspace and the trimming statistics are out of scope and stay with the reference.

`match_to(read)` is the reference's per-read call (a batch of one); `match_to_batch(reads)` is the
batched twin returning one record per read without creating Python objects.
"""
import numpy as np

from . import _abi, engine
from .align import (Match, START_WITHIN_SEQ1, START_WITHIN_SEQ2, STOP_WITHIN_SEQ1, STOP_WITHIN_SEQ2, SEMIGLOBAL,
                    _tuple_of)
from .util import IUPAC_BASES, RandomMatchProbability, expand_braces

# adapters/__init__.py:41-56
BACK = START_WITHIN_SEQ2 | STOP_WITHIN_SEQ2 | STOP_WITHIN_SEQ1
FRONT = START_WITHIN_SEQ2 | STOP_WITHIN_SEQ2 | START_WITHIN_SEQ1
PREFIX = STOP_WITHIN_SEQ2
SUFFIX = START_WITHIN_SEQ2
ANYWHERE = SEMIGLOBAL
LINKED = 'linked'


def _seq_of(read):
    return read if isinstance(read, str) else read.sequence


class Adapter(object):
    """Adapter(sequence, where, max_error_rate=0.1, min_overlap=3, read_wildcards=False,
    adapter_wildcards=True, name=None, indels=True, indel_cost=1, match_probability=None, max_rmp=None)
    -- adapters/__init__.py:259-322."""

    _counter = 0

    def __init__(self, sequence, where, max_error_rate=0.1, min_overlap=3, read_wildcards=False,
                 adapter_wildcards=True, name=None, indels=True, indel_cost=1, match_probability=None, max_rmp=None,
                 device=0):
        if len(sequence) == 0:
            raise ValueError("Empty adapter sequence")
        sequence = expand_braces(sequence.upper().replace('U', 'T'))     # adapters/__init__.py:267
        seq_set = set(sequence)
        if seq_set <= set('ACGT'):
            adapter_wildcards = False
        if adapter_wildcards and not seq_set <= IUPAC_BASES:
            raise ValueError("Invalid character(s) in adapter sequence: {}".format(','.join(seq_set - IUPAC_BASES)))
        if name is None:
            Adapter._counter += 1
            name = str(Adapter._counter)
        self.name = name
        self.sequence = sequence
        self.where = where
        self.max_error_rate = max_error_rate
        self.min_overlap = min(min_overlap, len(self.sequence))
        if max_rmp is not None and match_probability is None:
            match_probability = RandomMatchProbability()
        self.match_probability = match_probability
        self.max_rmp = max_rmp
        self.indels = indels
        self.indel_cost = indel_cost
        self.adapter_wildcards = adapter_wildcards
        self.read_wildcards = read_wildcards
        if where == ANYWHERE:
            self._front_flag = None
        else:
            self._front_flag = where not in (BACK, SUFFIX)
        self._device = device
        self._set = None

    def __len__(self):
        return len(self.sequence)

    def __repr__(self):
        return '<Adapter(name="{name}", sequence="{sequence}", where={where}, max_error_rate={max_error_rate}, '\
               'min_overlap={min_overlap}, read_wildcards={read_wildcards}, adapter_wildcards={adapter_wildcards}, '\
               'indels={indels})>'.format(**vars(self))

    def descriptor(self):
        """(AtrAdapterDesc, keepalive) with Adapter.match_to semantics switched on."""
        m = len(self.sequence)
        rmp_ok = None
        if self.max_rmp is not None:
            rmp_ok = np.zeros((m + 1, m + 1), dtype=np.uint8)
            for size in range(0, m + 1):
                for matches in range(0, size + 1):
                    rmp_ok[size, matches] = self.match_probability(matches, size) <= self.max_rmp
        return _abi.make_adapter_desc(
            self.sequence, self.max_error_rate, self.where, self.adapter_wildcards, self.read_wildcards,
            self.min_overlap, self.indel_cost if self.indels else 100000, match_to_semantics=True,
            no_indels=not self.indels, rmp_ok=rmp_ok)

    def _adapterset(self):
        key = engine.context_key(self._device)
        if self._set is None or self._set_key != key:
            self._set = engine.AdapterSet(engine.default_context(self._device), [self.descriptor()])
            self._set_key = key
        return self._set

    def match_to(self, read):
        """match_to(read) -> Match | None   (adapters/__init__.py:338-400)"""
        rec = self.match_to_batch([_seq_of(read)])[0]
        return self.match_from_record(rec, read)

    def match_to_batch(self, reads, win=None):
        """reads: list of str or (ascii, offsets). Returns a MATCH_DTYPE array (status 0 = None)."""
        ascii, offsets = reads if isinstance(reads, tuple) else engine.encode_reads(reads)
        return self._adapterset().locate_host(ascii, offsets, win=win, fold_case=True)

    def match_from_record(self, rec, read=None):
        st = int(rec["status"])
        if st == _abi.ATR_ST_NONE:
            return None
        if st == _abi.ATR_ST_INVALID:
            raise ValueError('A Match requires at least one matching position.')
        return Match(*_tuple_of(rec), front=self._front_flag, adapter=self, read=read)


class LinkedMatch(object):
    """adapters/__init__.py:613-635"""

    def __init__(self, front_match, back_match, adapter):
        self.front_match = front_match
        self.back_match = back_match
        self.adapter = adapter
        assert front_match is not None


class LinkedAdapter(object):
    """LinkedAdapter(front_sequence, back_sequence, front_anchored=True, back_anchored=False, **kwargs)
    -- adapters/__init__.py:637-690."""

    def __init__(self, front_sequence, back_sequence, front_anchored=True, back_anchored=False, name=None, **kwargs):
        assert front_anchored and not back_anchored
        self.front_anchored = front_anchored
        self.back_anchored = back_anchored
        self.where = LINKED
        self.name = name
        self.front_adapter = Adapter(front_sequence, where=PREFIX if front_anchored else FRONT, name=None, **kwargs)
        self.back_adapter = Adapter(back_sequence, where=SUFFIX if back_anchored else BACK, name=None, **kwargs)

    def match_to(self, read):
        seq = _seq_of(read)
        front_match = self.front_adapter.match_to(read)
        if front_match is None:
            return None
        rest = seq[front_match.rstop:] if isinstance(read, str) else read[front_match.rstop:]
        back_match = self.back_adapter.match_to(rest)
        return LinkedMatch(front_match, back_match, self)

    def match_to_batch(self, reads):
        """Returns (front MATCH_DTYPE array, back MATCH_DTYPE array). The back adapter is aligned inside the
        window [front.rstop, len) of the SAME packed read, so its coordinates are relative to
        read[front.rstop:] exactly like the reference's slice (:683-689); back status is NONE where the front
        adapter did not match."""
        ascii, offsets = reads if isinstance(reads, tuple) else engine.encode_reads(reads)
        front = self.front_adapter.match_to_batch((ascii, offsets))
        n = len(offsets) - 1
        lens = np.diff(offsets)
        win = np.zeros((n, 2), dtype=np.uint16)
        hit = front["status"] == _abi.ATR_ST_MATCH
        win[:, 0] = np.where(hit, front["rstop"], lens).astype(np.uint16)
        win[:, 1] = lens.astype(np.uint16)
        back = self.back_adapter.match_to_batch((ascii, offsets), win=win)
        back["status"][~hit] = _abi.ATR_ST_NONE
        return front, back
