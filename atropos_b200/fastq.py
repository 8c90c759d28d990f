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
from .adapters import ANYWHERE, BACK, LINKED, SUFFIX


class FormatError(Exception):
    """atropos.io.seqio.FormatError"""


_BASES = ('A', 'C', 'G', 'T', '')


def new_ops_stats():
    """Trimmer.trimmed_bases per modifier ([read 1, read 2]) and FilterWrapper.filtered per filter, as the report
    names them (modifiers.py:84-88, filters.py:48-52), plus the records written."""
    d = {"bp_cut": [0, 0], "bp_quality": [0, 0], "bp_n_ends": [0, 0], "bp_nextseq": [0, 0]}
    d.update({k: 0 for k in _abi.OPS_STAT_KEYS})
    return d


def merge_ops_stats(mine, theirs):
    for k, v in theirs.items():
        if isinstance(v, list):
            mine[k] = [a + b for a, b in zip(mine[k], v)]
        else:
            mine[k] += v


def _add_ops_stats(dst, st):
    for k in ("bp_cut", "bp_quality", "bp_n_ends", "bp_nextseq"):
        dst[k] = [dst[k][i] + int(getattr(st, k)[i]) for i in range(2)]
    for k in _abi.OPS_STAT_KEYS:
        dst[k] += int(getattr(st, k))


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
        self.ops = new_ops_stats()

    def merge(self, other):
        merge_ops_stats(self.ops, other.ops)
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

    if kind == _abi.ATR_FQ_MORE_IN_1:
        return "Reads are improperly paired. There are more reads in file 1 than in file 2."
    if kind == _abi.ATR_FQ_MORE_IN_2:
        return "Reads are improperly paired. There are more reads in file 2 than in file 1."
    if kind == _abi.ATR_FQ_CORRECTION:
        return "error correction would raise in the reference (reads of unequal length or bytes outside the complement table)"
    if kind == _abi.ATR_FQ_EMPTY_NAME:
        return "a read name without any token (the reference raises IndexError in sequence_names_match)"

    def content(b, e, t=None):
        s = bytes((text if t is None else t)[b:e])
        return (s[:-1] if s.endswith(b"\r") else s).decode("latin-1")

    if kind == _abi.ATR_FQ_PAIR_NAMES:      # io/seqio.py:448-452; text = (text1, text2)
        return "Reads are improperly paired. Read name '{0}' in file 1 does not match '{1}' in file 2.".format(
            content(err.line_begin, err.line_end, text[0])[1:], content(err.line_begin2, err.line_end2, text[1])[1:])
    if isinstance(text, tuple):
        text = text[err.file]

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

    def __init__(self, adapters, times=1, max_len=512, device=0, chunk_bytes=0, **read_ops):
        """read_ops: the command's modifier / filter options around the adapter stage, see _abi.make_read_ops (cut,
        quality_cutoff, quality_base, trim_n, minimum_length, maximum_length, max_n, discard_trimmed,
        discard_untrimmed), applied in the default operation order."""
        self.ops = _abi.make_read_ops(**read_ops)
        self.adapters = list(adapters)
        self.times = int(times)
        self.max_len = int(max_len)
        self.ctx = engine.default_context(device)
        self._back_set = None
        if len(self.adapters) == 1 and getattr(self.adapters[0], "where", None) == LINKED:
            # "-a FRONT...BACK": statistics index 0 = the front adapter, 1 = the back adapter
            linked = self.adapters[0]
            if self.times != 1:
                raise ValueError("a linked adapter works with times == 1 only")
            self.adapters = [linked.front_adapter, linked.back_adapter]
            self._set = engine.AdapterSet(self.ctx, [linked.front_adapter.descriptor()])
            self._back_set = engine.AdapterSet(self.ctx, [linked.back_adapter.descriptor()])
        else:
            if any(getattr(a, "where", None) == LINKED for a in self.adapters):
                raise ValueError("a linked adapter cannot be combined with other adapters (the reference fails there too)")
            self._set = engine.AdapterSet(self.ctx, [a.descriptor() for a in self.adapters])
        self.max_errors = max(int(a.max_error_rate * len(a.sequence)) for a in self.adapters)
        self.chunk_bytes = int(chunk_bytes)

    def new_stats(self):
        return TrimStats(len(self.adapters), self.max_len, self.max_errors)

    def trim(self, text, final=True, stats=None, out=None):
        """text: bytes / uint8 array (host). Returns (out uint8 array view, stats, consumed)."""
        buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
        n = int(buf.size)
        if out is None:
            out = np.empty(n + 1, dtype=np.uint8)          # an unterminated last record gains its newline
        if stats is None:
            stats = self.new_stats()
        opts = _abi.AtrTrimOpts(self.times, self.max_len, self.max_errors, int(bool(final)), self.chunk_bytes,
                                self._back_set.handle if self._back_set is not None else None, self.ops)
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
        _add_ops_stats(stats.ops, st.ops)
        if st.overflow:
            raise OverflowError("a removed length exceeds max_len=%d: create the FastqTrimmer with a larger max_len" % self.max_len)
        return out[:nout.value], stats, int(consumed.value)


def open_fastq(path, mode="rb", compresslevel=6):
    """Open a FASTQ file for the streaming helpers below by its suffix, like the reference's xopen
    (atropos/io/compression.py:17-71, io/__init__.py:128-173): .gz (multi-member files included), .bz2, .xz / .lzma, else
    plain. Decompression is host work (zlib); the text still crosses PCIe uncompressed."""
    p = path.decode() if isinstance(path, bytes) else path
    if p.endswith(".gz"):
        import gzip
        return gzip.open(p, mode, compresslevel=compresslevel) if "w" in mode else gzip.open(p, mode)
    if p.endswith(".bz2"):
        import bz2
        return bz2.open(p, mode)
    if p.endswith((".xz", ".lzma")):
        import lzma
        return lzma.open(p, mode)
    return open(p, mode)


def _stream_blocks(fh, block_bytes):
    while True:
        block = fh.read(block_bytes)
        yield block
        if not block:
            return


def trim_file(trimmer, src, dst, block_bytes=1 << 28):
    """Stream a FASTQ file through a FastqTrimmer block by block: the partial last record of a block is carried
    into the next one (`consumed`). src / dst: paths or binary file objects. Returns the accumulated TrimStats."""
    fin = open_fastq(src, "rb") if isinstance(src, (str, bytes)) else src
    fout = open_fastq(dst, "wb") if isinstance(dst, (str, bytes)) else dst
    try:
        stats, carry = trimmer.new_stats(), b""
        for block in _stream_blocks(fin, block_bytes):
            text = carry + block
            out, stats, consumed = trimmer.trim(text, final=not block, stats=stats)
            fout.write(out.tobytes())
            carry = text[consumed:]
        return stats
    finally:
        if fin is not src:
            fin.close()
        if fout is not dst:
            fout.close()


def trim_file_pair(trimmer, src1, src2, dst1, dst2, block_bytes=1 << 27, dst_merged=None):
    """Paired-end twin of trim_file for a FastqPairTrimmer: the two inputs advance independently.
    dst_merged: the --merged-output file of a trimmer created with merge_overlapping=True."""
    opened = []

    def _open(x, mode):
        if isinstance(x, (str, bytes)):
            fh = open_fastq(x, mode)
            opened.append(fh)
            return fh
        return x

    f1, f2, o1, o2 = _open(src1, "rb"), _open(src2, "rb"), _open(dst1, "wb"), _open(dst2, "wb")
    om = _open(dst_merged, "wb") if dst_merged is not None else None
    try:
        stats, c1, c2 = trimmer.new_stats(), b"", b""
        eof1 = eof2 = False
        while True:
            b1 = b"" if eof1 else f1.read(block_bytes)
            b2 = b"" if eof2 else f2.read(block_bytes)
            eof1, eof2 = eof1 or not b1, eof2 or not b2
            t1, t2 = c1 + b1, c2 + b2
            final = eof1 and eof2
            outs, stats, consumed = trimmer.trim(t1, t2, final=final, stats=stats)
            o1.write(outs[0].tobytes())
            o2.write(outs[1].tobytes())
            if om is not None and len(outs) > 2:
                om.write(outs[2].tobytes())
            c1, c2 = t1[consumed[0]:], t2[consumed[1]:]
            if final:
                return stats
    finally:
        for fh in opened:
            fh.close()


class PairTrimStats(object):
    """InsertAdapterCutter.summarize() (commands/trim/modifiers.py:498-509) or, in adapter mode, the two AdapterCutters'
    (:189-195): per read and adapter the statistics Adapter.trimmed() keeps."""

    def __init__(self, max_len, max_errors, n_adapters=(1, 1)):
        self.max_len, self.max_errors, self.n_adapters = max_len, max_errors, tuple(max(1, k) for k in n_adapters)
        shape = lambda i: (self.n_adapters[i], max_len + 1, max_errors + 1)
        self.errors_back = [np.zeros(shape(i), dtype=np.int64) for i in range(2)]
        self.errors_front = [np.zeros(shape(i), dtype=np.int64) for i in range(2)]
        self.adjacent = [np.zeros((self.n_adapters[i], 5), dtype=np.int64) for i in range(2)]
        self.records = self.insert_matches = self.overflow = 0
        self.with_adapters, self.bp_in, self.bp_out = [0, 0], [0, 0], [0, 0]
        self.records_corrected, self.bp_corrected = 0, [0, 0]      # ErrorCorrectorMixin.summarize (modifiers.py:352-357)
        # MergeOverlapping + MergedReadFilter (FastqPairTrimmer(merge_overlapping=True)): pairs merged, merged reads written
        # (and their bases: the report adds them to bp_written[0]), and the merge stage's own correction counters
        self.merged = self.merged_written = self.bp_merged_written = 0
        self.merge_records_corrected, self.merge_bp_corrected = 0, [0, 0]
        self.ops = new_ops_stats()

    def merge(self, other):
        merge_ops_stats(self.ops, other.ops)
        self.records_corrected += other.records_corrected
        self.merged += other.merged
        self.merged_written += other.merged_written
        self.bp_merged_written += other.bp_merged_written
        self.merge_records_corrected += other.merge_records_corrected
        for i in range(2):
            self.bp_corrected[i] += other.bp_corrected[i]
            self.merge_bp_corrected[i] += other.merge_bp_corrected[i]
            self.errors_back[i] += other.errors_back[i]
            self.errors_front[i] += other.errors_front[i]
            self.adjacent[i] += other.adjacent[i]
            self.with_adapters[i] += other.with_adapters[i]
            self.bp_in[i] += other.bp_in[i]
            self.bp_out[i] += other.bp_out[i]
        self.records += other.records
        self.insert_matches += other.insert_matches
        self.overflow += other.overflow
        return self

    def adapter_summary(self, i, a=0, where=BACK):
        """statistics of adapter `a` of read `i` (0 / 1), the keys Adapter.summarize() emits for its type"""
        d = {}
        if where not in (BACK, SUFFIX):
            d["errors_front"] = TrimStats._nested(self.errors_front[i][a])
            d["lengths_front"] = {ln: sum(v.values()) for ln, v in d["errors_front"].items()}
        if where in (ANYWHERE, BACK, SUFFIX):
            d["errors_back"] = TrimStats._nested(self.errors_back[i][a])
            d["lengths_back"] = {ln: sum(v.values()) for ln, v in d["errors_back"].items()}
        if where in (BACK, SUFFIX):
            d["adjacent_bases"] = {b: int(self.adjacent[i][a, k]) for k, b in enumerate(_BASES)}
        return d


class FastqPairTrimmer(object):
    """Paired-end twin of FastqTrimmer: the reference's `--aligner insert` pipeline for two FASTQ texts read in
    lockstep (PairedSequenceReader io/seqio.py:397-453 -> InsertAdapterCutter modifiers.py:359-496 -> two
    FastqFormat outputs) in one `atr_trim_fastq_pe_host` call.

    adapter1 / adapter2: atropos_b200.adapters.Adapter, the 3' adapters of read 1 / read 2 as the reference's command
    line builds them for insert mode (max_rmp 1e-6, min_overlap 1, indel_cost 3: trim/cli.py:667-679);
    insert_aligner: atropos_b200.align.InsertAligner with the same sequences."""

    def __init__(self, adapter1, adapter2, insert_aligner=None, symmetric=True, min_insert_overlap=1, max_len=256, device=0,
                 chunk_bytes=0, times=1, mismatch_action=None, merge_overlapping=False, merge_min_overlap=0.9,
                 merge_error_rate=0.2, merged_output=True, **read_ops):
        """mismatch_action: --correct-mismatches ('liberal', 'conservative', 'N'; insert mode, or with merge_overlapping).
        merge_overlapping: --merge-overlapping with --merge-min-overlap / --merge-error-rate (MergeOverlapping as the last
        modifier, MergedReadFilter as the first filter: commands/trim/__init__.py:546-552, :576-579); trim() then returns
        a third text, the merged reads (--merged-output). merged_output=False: no such file, merged pairs are discarded.
        insert_aligner given: `--aligner insert` (adapter1 / adapter2 = the one 3' adapter of each read).
        insert_aligner None: the command's default `--aligner adapter`: adapter1 / adapter2 are lists of Adapters (or
        None) for read 1 / read 2, cut independently with `times` rounds each (commands/trim/__init__.py:457-476)."""
        self.ops = _abi.make_read_ops(**read_ops)
        self.aligner = insert_aligner
        as_list = lambda a: [] if a is None else (list(a) if isinstance(a, (list, tuple)) else [a])
        self.adapters = [as_list(adapter1), as_list(adapter2)]
        if insert_aligner is not None:
            for ads in self.adapters:
                if len(ads) != 1 or ads[0].where != BACK:
                    raise ValueError("Insert aligner requires a single 3' adapter for each read")
        self.adapter1 = self.adapters[0][0] if self.adapters[0] else None
        self.adapter2 = self.adapters[1][0] if self.adapters[1] else None
        self.symmetric, self.min_insert_overlap, self.times = bool(symmetric), int(min_insert_overlap), int(times)
        self.mismatch_action = _abi.MISMATCH_ACTIONS[mismatch_action]
        self.merge = (float(merge_min_overlap), float(merge_error_rate)) if merge_overlapping else None
        self.merged_output = bool(merged_output)
        if self.mismatch_action and insert_aligner is None and self.merge is None:
            raise ValueError("error correction needs the insert aligner or merge_overlapping")
        self.max_len = int(max_len)
        every = self.adapters[0] + self.adapters[1]
        if insert_aligner is not None:
            self.max_errors = max(len(a.sequence) for a in every)
        else:
            self.max_errors = max([int(a.max_error_rate * len(a.sequence)) for a in every] or [0])
        self.ctx = engine.default_context(device)
        self._sets = [engine.AdapterSet(self.ctx, [a.descriptor() for a in ads]) if ads else None for ads in self.adapters]
        self._iset = insert_aligner._insertset(self.max_len) if insert_aligner is not None else None
        self.chunk_bytes = int(chunk_bytes)

    def new_stats(self):
        return PairTrimStats(self.max_len, self.max_errors, (len(self.adapters[0]), len(self.adapters[1])))

    def trim(self, text1, text2, final=True, stats=None, out1=None, out2=None, out_merged=None):
        """Returns ((out1, out2) uint8 views, stats, (consumed1, consumed2)); with merge_overlapping the first element is
        (out1, out2, merged)."""
        b1 = np.frombuffer(text1, dtype=np.uint8) if not isinstance(text1, np.ndarray) else text1
        b2 = np.frombuffer(text2, dtype=np.uint8) if not isinstance(text2, np.ndarray) else text2
        if out1 is None:
            out1 = np.empty(int(b1.size) + 1, dtype=np.uint8)
        if out2 is None:
            out2 = np.empty(int(b2.size) + 1, dtype=np.uint8)
        if stats is None:
            stats = self.new_stats()
        opts = _abi.AtrTrimPeOpts(int(self.symmetric), self.min_insert_overlap, self.max_len, self.max_errors,
                                  int(bool(final)), self.times, self.mismatch_action, 0, self.chunk_bytes, self.ops)
        st = _abi.AtrTrimPeStats()
        for i in range(2):
            st.errors_back[i] = stats.errors_back[i].ctypes.data
            st.errors_front[i] = stats.errors_front[i].ctypes.data
            st.adjacent_bases[i] = stats.adjacent[i].ctypes.data
        err = _abi.AtrFastqError()
        nout, consumed = (C.c_int64 * 2)(), (C.c_int64 * 2)()
        L = _lib.load()
        h = lambda x: x.handle if x is not None else None
        if self.merge is None:
            rc = L.atr_trim_fastq_pe_host(self.ctx.handle, h(self._iset), h(self._sets[0]), h(self._sets[1]), C.byref(opts),
                                          b1.ctypes.data if b1.size else None, int(b1.size),
                                          b2.ctypes.data if b2.size else None, int(b2.size),
                                          out1.ctypes.data, int(out1.size), out2.ctypes.data, int(out2.size), nout, consumed,
                                          C.byref(st), C.byref(err))
        else:
            mo, ms = _abi.AtrMergeOpts(*self.merge), _abi.AtrMergeStats()
            nout = (C.c_int64 * 3)()
            if out_merged is None and self.merged_output:
                out_merged = np.empty(int(b1.size) + int(b2.size) + 1, dtype=np.uint8)
            rc = L.atr_trim_fastq_pe_merge_host(self.ctx.handle, h(self._iset), h(self._sets[0]), h(self._sets[1]), C.byref(opts),
                                                C.byref(mo), b1.ctypes.data if b1.size else None, int(b1.size),
                                                b2.ctypes.data if b2.size else None, int(b2.size),
                                                out1.ctypes.data, int(out1.size), out2.ctypes.data, int(out2.size),
                                                out_merged.ctypes.data if self.merged_output else None,
                                                int(out_merged.size) if self.merged_output else 0, nout, consumed,
                                                C.byref(st), C.byref(ms), C.byref(err))
        if rc == _abi.ATR_E_FORMAT:
            raise FormatError(format_error_message((b1, b2), err))
        _lib.check(rc, self.ctx.handle)
        stats.records += int(st.records)
        stats.insert_matches += int(st.insert_matches)
        stats.overflow += int(st.overflow)
        stats.records_corrected += int(st.records_corrected)
        for i in range(2):
            stats.bp_corrected[i] += int(st.bp_corrected[i])
            stats.with_adapters[i] += int(st.with_adapters[i])
            stats.bp_in[i] += int(st.bp_in[i])
            stats.bp_out[i] += int(st.bp_out[i])
        _add_ops_stats(stats.ops, st.ops)
        if st.overflow:
            raise OverflowError("a removed length exceeds max_len=%d: create the FastqPairTrimmer with a larger max_len" % self.max_len)
        if self.merge is not None:
            stats.merged += int(ms.merged)
            stats.merged_written += int(ms.merged_written)
            stats.bp_merged_written += int(ms.bp_merged_written)
            stats.merge_records_corrected += int(ms.records_corrected)
            for i in range(2):
                stats.merge_bp_corrected[i] += int(ms.bp_corrected[i])
            merged = out_merged[:nout[2]] if self.merged_output else np.empty(0, dtype=np.uint8)
            return (out1[:nout[0]], out2[:nout[1]], merged), stats, (int(consumed[0]), int(consumed[1]))
        return (out1[:nout[0]], out2[:nout[1]]), stats, (int(consumed[0]), int(consumed[1]))
