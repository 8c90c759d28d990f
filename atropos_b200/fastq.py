"""FASTQ text in -> trimmed FASTQ text out on the GPU: the batch-staged twin of the reference's per-record pipeline

    FastqReader.__iter__      atropos/io/_seqio.pyx:180-245
    AdapterCutter.__call__    atropos/commands/trim/modifiers.py:124-187   (action='trim', `times` rounds)
    Adapter.trimmed           atropos/adapters/__init__.py:413-436        (statistics of the report)
    FastqFormat.format        atropos/io/seqio.py:686-700

for single-end reads ("next" rows f-1/f-2/f-3 of SURVEY.md section 8). One call = `atr_trim_fastq_host`
(include/atropos_b200.h): the text crosses PCIe once in each direction, everything between happens in HBM.
Malformed input raises `FormatError` with the reference's own message.
"""
import ctypes as C

import numpy as np

from . import _abi, _lib, engine
from .adapters import ANYWHERE, BACK, SUFFIX


class FormatError(Exception):
    """atropos.io.seqio.FormatError"""


_BASES = ('A', 'C', 'G', 'T', '')


class TrimStats(object):
    """What the reference's report holds for one AdapterCutter (commands/trim/modifiers.py:189-195 and
    adapters/__init__.py:474-505), accumulated over calls / shards with `merge`."""

    def __init__(self, n_adapters, max_len, max_errors):
        self.n_adapters, self.max_len, self.max_errors = n_adapters, max_len, max_errors
        shape = (n_adapters, max_len + 1, max_errors + 1)
        self.errors_front = np.zeros(shape, dtype=np.int64)
        self.errors_back = np.zeros(shape, dtype=np.int64)
        self.adjacent = np.zeros((n_adapters, 5), dtype=np.int64)
        self.records = self.with_adapters = self.bp_in = self.bp_out = self.overflow = 0

    def merge(self, other):
        self.errors_front += other.errors_front
        self.errors_back += other.errors_back
        self.adjacent += other.adjacent
        for k in ("records", "with_adapters", "bp_in", "bp_out", "overflow"):
            setattr(self, k, getattr(self, k) + getattr(other, k))
        return self

    @staticmethod
    def _nested(hist):
        out = {}
        for ln, e in zip(*np.nonzero(hist)):
            out.setdefault(int(ln), {})[int(e)] = int(hist[ln, e])
        return out

    def adapter_summary(self, a, where):
        """lengths_* / errors_* / adjacent_bases of adapter `a` as the reference's dicts (only the keys
        Adapter.summarize() emits for this adapter type, adapters/__init__.py:493-503)."""
        d = {}
        if where not in (BACK, SUFFIX):
            d["errors_front"] = self._nested(self.errors_front[a])
            d["lengths_front"] = {ln: sum(v.values()) for ln, v in d["errors_front"].items()}
        if where in (ANYWHERE, BACK, SUFFIX):
            d["errors_back"] = self._nested(self.errors_back[a])
            d["lengths_back"] = {ln: sum(v.values()) for ln, v in d["errors_back"].items()}
        if where in (BACK, SUFFIX):
            d["adjacent_bases"] = {b: int(self.adjacent[a, i]) for i, b in enumerate(_BASES)}
        return d


def format_error_message(text, err):
    """The message FastqReader raises for this atr_fastq_error (io/_seqio.pyx:192-245). `text`: the bytes that
    were passed in."""
    kind = err.kind
    if kind == _abi.ATR_FQ_TRUNCATED:
        return "FASTQ file ended prematurely"
    if kind == _abi.ATR_FQ_LENGTH:
        return "Error creating sequence record at line 4"
    if kind == _abi.ATR_FQ_BARE_CR:
        return "carriage return without newline: not supported by the GPU FASTQ path"
    if kind == _abi.ATR_FQ_TOO_LONG:
        return "FASTQ record outside the supported sizes (read > 32767 nt, header > 65535 bytes or record > chunk)"
    if kind == _abi.ATR_FQ_INVALID_MATCH:
        return "A Match requires at least one matching position."

    def content(b, e):
        s = bytes(text[b:e])
        return (s[:-1] if s.endswith(b"\r") else s).decode("latin-1")

    line = content(err.line_begin, err.line_end)
    if kind == _abi.ATR_FQ_NO_AT:           # the raw line, newline included (universal newlines: "\n")
        raw = line + ("\n" if err.terminated else "")
        return "Line 1 in FASTQ file is expected to start with '@', but found {0!r}".format(raw[:10])
    sliced = line if err.terminated else line[:-1]
    if kind == _abi.ATR_FQ_NO_PLUS:
        return "Line 3 in FASTQ file is expected to start with '+', but found {0!r}".format(sliced[:10])
    if kind == _abi.ATR_FQ_NAME_MISMATCH:   # the header is two lines up
        head = bytes(text[:max(err.line_begin - 1, 0)])                          # up to the newline ending the sequence line
        hdr_e = head.rfind(b"\n")                                                 # newline ending the header line
        hdr_b = head.rfind(b"\n", 0, max(hdr_e, 0)) + 1
        name = content(hdr_b, hdr_e)[1:]
        return ("At line 3: Sequence descriptions in the FASTQ file don't match ({0!r} != {1!r}).\n"
                "The second sequence description must be either empty or equal to the first "
                "description.".format(name, sliced[1:]))
    return "malformed FASTQ (kind %d)" % kind


class FastqTrimmer(object):
    """trimmer = FastqTrimmer([Adapter(...), ...], times=1); out_bytes, stats = trimmer.trim(text_bytes)

    `adapters`: atropos_b200.adapters.Adapter objects in the order the reference's AdapterCutter would try them
    (the command line collects -a, then -b, then -g). Linked adapters are not handled by this path."""

    def __init__(self, adapters, times=1, max_len=512, device=0, chunk_bytes=0):
        self.adapters = list(adapters)
        self.times = int(times)
        self.max_len = int(max_len)
        self.max_errors = max(int(a.max_error_rate * len(a.sequence)) for a in self.adapters)
        self.ctx = engine.default_context(device)
        self._set = engine.AdapterSet(self.ctx, [a.descriptor() for a in self.adapters])
        self.chunk_bytes = int(chunk_bytes)

    def new_stats(self):
        return TrimStats(len(self.adapters), self.max_len, self.max_errors)

    def trim(self, text, final=True, stats=None, out=None):
        """text: bytes / uint8 array (host). Returns (out uint8 array view, stats, consumed)."""
        buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
        n = int(buf.size)
        if out is None:
            out = np.empty(max(n, 1), dtype=np.uint8)
        if stats is None:
            stats = self.new_stats()
        opts = _abi.AtrTrimOpts(self.times, self.max_len, self.max_errors, int(bool(final)), self.chunk_bytes)
        st = _abi.AtrTrimStats()
        st.errors_front = stats.errors_front.ctypes.data
        st.errors_back = stats.errors_back.ctypes.data
        st.adjacent_bases = stats.adjacent.ctypes.data
        err = _abi.AtrFastqError()
        nout, consumed = C.c_int64(0), C.c_int64(0)
        L = _lib.load()
        rc = L.atr_trim_fastq_host(self.ctx.handle, self._set.handle, C.byref(opts), buf.ctypes.data if n else None, n,
                                   out.ctypes.data, int(out.size), C.byref(nout), C.byref(consumed), C.byref(st), C.byref(err))
        if rc == _abi.ATR_E_FORMAT:
            raise FormatError(format_error_message(buf, err))
        _lib.check(rc, self.ctx.handle)
        for k in ("records", "with_adapters", "bp_in", "bp_out", "overflow"):
            setattr(stats, k, getattr(stats, k) + int(getattr(st, k)))
        if st.overflow:
            raise OverflowError("a removed length exceeds max_len=%d: create the FastqTrimmer with a larger max_len" % self.max_len)
        return out[:nout.value], stats, int(consumed.value)
