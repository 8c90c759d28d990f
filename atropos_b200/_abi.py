"""ctypes mirror of include/atropos_b200.h (structures and constants only)."""
import ctypes as C

import numpy as np

ATR_ABI_VERSION = 1
ATR_OK, ATR_E_ARG, ATR_E_CUDA, ATR_E_NOMEM, ATR_E_LIMIT, ATR_E_FORMAT = 0, -1, -2, -3, -4, -5
(ATR_FQ_OK, ATR_FQ_NO_AT, ATR_FQ_NO_PLUS, ATR_FQ_NAME_MISMATCH, ATR_FQ_LENGTH, ATR_FQ_TRUNCATED, ATR_FQ_BARE_CR,
 ATR_FQ_TOO_LONG, ATR_FQ_INVALID_MATCH, ATR_FQ_MORE_IN_1, ATR_FQ_MORE_IN_2, ATR_FQ_PAIR_NAMES, ATR_FQ_EMPTY_NAME,
 ATR_FQ_CORRECTION) = range(14)
MISMATCH_ACTIONS = {None: 0, "liberal": 1, "conservative": 2, "N": 3}
ATR_ST_NONE, ATR_ST_MATCH, ATR_ST_ESCAPED, ATR_ST_INVALID, ATR_ST_KEYERROR = 0, 1, 2, 3, 4


class AtrMatch(C.Structure):
    _fields_ = [("astart", C.c_uint16), ("astop", C.c_uint16), ("rstart", C.c_uint16), ("rstop", C.c_uint16),
                ("matches", C.c_uint16), ("errors", C.c_uint16), ("adapter", C.c_int16), ("status", C.c_uint16)]


#: numpy view of an array of atr_match
MATCH_DTYPE = np.dtype([("astart", "<u2"), ("astop", "<u2"), ("rstart", "<u2"), ("rstop", "<u2"),
                        ("matches", "<u2"), ("errors", "<u2"), ("adapter", "<i2"), ("status", "<u2")])
#: numpy view of an array of atr_insert_result
INSERT_DTYPE = np.dtype([("insert", MATCH_DTYPE), ("match1", MATCH_DTYPE), ("match2", MATCH_DTYPE)])

assert C.sizeof(AtrMatch) == 16 and MATCH_DTYPE.itemsize == 16 and INSERT_DTYPE.itemsize == 48


class AtrMergeResult(C.Structure):
    _fields_ = [("r2_start", C.c_uint16), ("r2_stop", C.c_uint16), ("r1_start", C.c_uint16), ("r1_stop", C.c_uint16),
                ("matches", C.c_uint16), ("errors", C.c_uint16), ("min_overlap", C.c_uint16), ("status", C.c_uint8),
                ("action", C.c_uint8)]


#: numpy view of an array of atr_merge_result
MERGE_DTYPE = np.dtype([("r2_start", "<u2"), ("r2_stop", "<u2"), ("r1_start", "<u2"), ("r1_stop", "<u2"),
                        ("matches", "<u2"), ("errors", "<u2"), ("min_overlap", "<u2"), ("status", "u1"), ("action", "u1")])
ATR_MERGE_KEEP1, ATR_MERGE_TAKE2, ATR_MERGE_APPEND, ATR_MERGE_PREPEND = 1, 2, 3, 4

assert C.sizeof(AtrMergeResult) == 16 and MERGE_DTYPE.itemsize == 16


class AtrAdapterDesc(C.Structure):
    _fields_ = [("sequence", C.c_char_p), ("length", C.c_int32), ("max_error_rate", C.c_double),
                ("flags", C.c_int32), ("wildcard_ref", C.c_int32), ("wildcard_query", C.c_int32),
                ("min_overlap", C.c_int32), ("indel_cost", C.c_int32), ("match_to_semantics", C.c_int32),
                ("no_indels", C.c_int32), ("rmp_ok", C.c_void_p)]


class AtrInsertDesc(C.Structure):
    _fields_ = [("adapter1", C.c_char_p), ("adapter1_len", C.c_int32),
                ("adapter2", C.c_char_p), ("adapter2_len", C.c_int32),
                ("insert_max_rmp", C.c_double), ("adapter_max_rmp", C.c_double),
                ("min_insert_overlap", C.c_int32), ("max_insert_mismatch_frac", C.c_double),
                ("min_adapter_overlap", C.c_int32), ("max_adapter_mismatch_frac", C.c_double),
                ("adapter_check_cutoff", C.c_int32), ("adapter_wildcards", C.c_int32),
                ("read_wildcards", C.c_int32), ("max_len", C.c_int32),
                ("insert_prob", C.c_void_p), ("adapter_prob", C.c_void_p)]


class AtrFastqError(C.Structure):
    _fields_ = [("kind", C.c_int32), ("line_in_record", C.c_int32), ("record", C.c_int64), ("line_begin", C.c_int64),
                ("line_end", C.c_int64), ("terminated", C.c_int32), ("file", C.c_int32), ("line_begin2", C.c_int64),
                ("line_end2", C.c_int64)]


class AtrReadOps(C.Structure):
    _fields_ = [("cut_front", C.c_int32 * 2), ("cut_back", C.c_int32 * 2), ("quality_front", C.c_int32),
                ("quality_back", C.c_int32), ("quality_base", C.c_int32), ("trim_n", C.c_int32),
                ("minimum_length", C.c_int32), ("maximum_length", C.c_int32), ("discard_trimmed", C.c_int32),
                ("discard_untrimmed", C.c_int32), ("nextseq_trim", C.c_int32 * 2), ("legacy_first", C.c_int32),
                ("pair_filter_both", C.c_int32), ("max_n", C.c_double)]


class AtrReadOpsStats(C.Structure):
    _fields_ = [("bp_cut", C.c_int64 * 2), ("bp_quality", C.c_int64 * 2), ("bp_n_ends", C.c_int64 * 2),
                ("bp_nextseq", C.c_int64 * 2), ("too_short", C.c_int64), ("too_long", C.c_int64), ("too_many_n", C.c_int64),
                ("discarded_trimmed", C.c_int64), ("discarded_untrimmed", C.c_int64), ("records_written", C.c_int64)]


OPS_STAT_KEYS = ("too_short", "too_long", "too_many_n", "discarded_trimmed", "discarded_untrimmed", "records_written")


def make_read_ops(cut=(), cut2=(), quality_cutoff=None, quality_base=33, trim_n=False, minimum_length=None,
                  maximum_length=None, max_n=None, discard_trimmed=False, discard_untrimmed=False, legacy_first=False,
                  nextseq_trim=None, pair_filter="any"):
    """The `trim` command's options (trim/cli.py) -> atr_read_ops. cut / cut2: the -u / -U values (lists of ints);
    quality_cutoff: -q as the command normalises it, [back] or [front, back] (trim/cli.py:750-754)."""
    o = AtrReadOps()
    for i, lengths in enumerate((cut, cut2)):
        lengths = [lengths] if isinstance(lengths, int) else list(lengths or ())
        o.cut_front[i] = sum(x for x in lengths if x > 0)          # UnconditionalCutter.__init__ (modifiers.py:578-582)
        o.cut_back[i] = sum(x for x in lengths if x < 0)
    q = quality_cutoff
    if q is not None:
        q = [q] if isinstance(q, int) else list(q)
        if all(c <= 0 for c in q):
            q = None
        elif len(q) == 1:
            q = [0] + q
    o.quality_front, o.quality_back = (q[0], q[1]) if q else (0, 0)
    o.quality_base = int(quality_base)
    o.trim_n = int(bool(trim_n))
    o.minimum_length = int(minimum_length) if minimum_length and minimum_length > 0 else 0
    o.maximum_length = int(maximum_length) if maximum_length is not None else -1
    o.max_n = float(max_n) if max_n is not None else -1.0
    o.discard_trimmed, o.discard_untrimmed = int(bool(discard_trimmed)), int(bool(discard_untrimmed))
    # paired-end legacy mode (trim/cli.py:629-645): nothing on the command line touches read 2 -> filters see read 1 only
    o.legacy_first = int(bool(legacy_first))
    if pair_filter not in ("any", "both"):
        raise ValueError("pair_filter must be 'any' or 'both'")
    o.pair_filter_both = int(pair_filter == "both")      # --pair-filter: PairedWrapper.min_affected (trim/__init__.py:556-557)
    # --nextseq-trim: NextseqQualityTrimmer on both reads, on read 1 only in legacy mode (PairedEndModifiers.add_modifier)
    o.nextseq_trim[0] = -1 if nextseq_trim is None else int(nextseq_trim)
    o.nextseq_trim[1] = -1 if (nextseq_trim is None or legacy_first) else int(nextseq_trim)
    return o


class AtrTrimOpts(C.Structure):
    _fields_ = [("times", C.c_int32), ("max_len", C.c_int32), ("max_errors", C.c_int32), ("final_chunk", C.c_int32),
                ("chunk_bytes", C.c_int64), ("linked_back", C.c_void_p), ("ops", AtrReadOps)]


class AtrTrimStats(C.Structure):
    _fields_ = [("records", C.c_int64), ("with_adapters", C.c_int64), ("bp_in", C.c_int64), ("bp_out", C.c_int64),
                ("overflow", C.c_int64), ("errors_front", C.c_void_p), ("errors_back", C.c_void_p),
                ("adjacent_bases", C.c_void_p), ("ops", AtrReadOpsStats)]


class AtrTrimPeOpts(C.Structure):
    _fields_ = [("symmetric", C.c_int32), ("min_insert_overlap", C.c_int32), ("max_len", C.c_int32),
                ("max_errors", C.c_int32), ("final_chunk", C.c_int32), ("times", C.c_int32), ("mismatch_action", C.c_int32), ("pad", C.c_int32),
                ("chunk_bytes", C.c_int64),
                ("ops", AtrReadOps)]


class AtrMergeOpts(C.Structure):
    _fields_ = [("min_overlap", C.c_double), ("error_rate", C.c_double)]


class AtrMergeStats(C.Structure):
    _fields_ = [("merged", C.c_int64), ("merged_written", C.c_int64), ("bp_merged_written", C.c_int64),
                ("records_corrected", C.c_int64), ("bp_corrected", C.c_int64 * 2)]


class AtrTrimPeStats(C.Structure):
    _fields_ = [("records", C.c_int64), ("insert_matches", C.c_int64), ("with_adapters", C.c_int64 * 2),
                ("bp_in", C.c_int64 * 2), ("bp_out", C.c_int64 * 2), ("overflow", C.c_int64),
                ("errors_back", C.c_void_p * 2), ("adjacent_bases", C.c_void_p * 2), ("errors_front", C.c_void_p * 2),
                ("records_corrected", C.c_int64), ("bp_corrected", C.c_int64 * 2), ("ops", AtrReadOpsStats)]


def make_adapter_desc(sequence, max_error_rate, flags, wildcard_ref=False, wildcard_query=False, min_overlap=1,
                      indel_cost=1, match_to_semantics=False, no_indels=False, rmp_ok=None):
    """Build an AtrAdapterDesc; returns (desc, keepalive) -- keep `keepalive` referenced while desc is used."""
    seq = sequence if isinstance(sequence, bytes) else sequence.encode("ascii")
    d = AtrAdapterDesc()
    d.sequence = seq
    d.length = len(seq)
    d.max_error_rate = float(max_error_rate)
    d.flags = int(flags)
    d.wildcard_ref = int(bool(wildcard_ref))
    d.wildcard_query = int(bool(wildcard_query))
    d.min_overlap = int(min_overlap)
    d.indel_cost = int(indel_cost)
    d.match_to_semantics = int(bool(match_to_semantics))
    d.no_indels = int(bool(no_indels))
    keep = [seq]
    if rmp_ok is not None:
        tab = np.ascontiguousarray(rmp_ok, dtype=np.uint8)
        assert tab.size == (len(seq) + 1) ** 2
        d.rmp_ok = tab.ctypes.data
        keep.append(tab)
    else:
        d.rmp_ok = None
    return d, keep
