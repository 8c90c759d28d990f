"""Build + ctypes access to tests/host_sim/libsim.so (CPU run of the per-read device functions)."""
import ctypes as C
import os
import subprocess

from atropos_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_sim", "sim.cpp")
LIB = os.path.join(HERE, "host_sim", "libsim.so")
CSRC = os.path.join(os.path.dirname(HERE), "atropos_b200", "csrc")
_lib = None


class SimOpsCounters(C.Structure):          # FqOpsCounters (fastq_core.cuh): same fields as atr_read_ops_stats, unsigned
    _fields_ = [(n, t) for n, t in _abi.AtrReadOpsStats._fields_]


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-o", LIB, SRC])
        L = C.CDLL(LIB)
        L.sim_locate.argtypes = [C.POINTER(_abi.AtrAdapterDesc), C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.POINTER(_abi.AtrMatch), C.POINTER(C.c_int)]
        L.sim_locate.restype = C.c_int
        L.sim_match_insert.argtypes = [C.POINTER(_abi.AtrInsertDesc), C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                       C.c_void_p, C.POINTER(C.c_int)]
        L.sim_match_insert.restype = C.c_int
        L.sim_multi_locate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_int)]
        L.sim_multi_locate.restype = C.c_int
        L.sim_filter_compare.argtypes = [C.POINTER(_abi.AtrAdapterDesc), C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_int)]
        L.sim_filter_compare.restype = C.c_int
        L.sim_set_qg.argtypes = [C.c_int]
        L.sim_set_qg.restype = None
        L.sim_compare_prefixes.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.sim_compare_prefixes.restype = C.c_int
        L.sim_trim_fastq.argtypes = [C.POINTER(_abi.AtrAdapterDesc), C.c_int, C.POINTER(_abi.AtrTrimOpts), C.c_char_p,
                                     C.c_longlong, C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_abi.AtrFastqError),
                                     C.POINTER(SimOpsCounters), C.POINTER(_abi.AtrAdapterDesc)]
        L.sim_trim_fastq.restype = C.c_int
        L.sim_trim_fastq_pe.argtypes = [C.POINTER(_abi.AtrInsertDesc), C.POINTER(_abi.AtrAdapterDesc), C.c_int,
                                        C.POINTER(_abi.AtrAdapterDesc), C.c_int, C.c_void_p, C.c_void_p,
                                        C.POINTER(_abi.AtrTrimPeOpts), C.c_char_p,
                                        C.c_longlong, C.c_char_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_abi.AtrFastqError),
                                        C.POINTER(SimOpsCounters), C.POINTER(C.c_longlong), C.POINTER(_abi.AtrMergeOpts),
                                        C.c_void_p, C.POINTER(C.c_longlong)]
        L.sim_trim_fastq_pe.restype = C.c_int
        L.sim_merge_overlap.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                        C.POINTER(_abi.AtrMergeResult)]
        L.sim_merge_overlap.restype = C.c_int
        _lib = L
    return _lib


def filter_compare(read, desc, lo=0, hi=None, fold_case=False):
    """First funnel stage both ways (Shift-And automaton / q-gram sampling) for one read: None if the adapter has no
    q-gram form, else (automaton SaResult 6-tuple, q-gram SaResult 6-tuple, (hmin, hmax) x 2, need_tail x 2, step)."""
    rb = read if isinstance(read, bytes) else read.encode("latin-1")
    out = (C.c_int * 24)()
    if not lib().sim_filter_compare(C.byref(desc), rb, len(rb), lo, len(rb) if hi is None else hi, int(fold_case), out):
        return None
    o = list(out)
    return tuple(o[0:6]), tuple(o[6:12]), (o[12], o[13]), (o[14], o[15]), o[16], o[17], o[18]


def merge_overlap(seq1, seq2, insert_matched, min_overlap, error_rate):
    """AtrMergeResult of one pair (merge_core.cuh on the CPU)."""
    b1 = seq1 if isinstance(seq1, bytes) else seq1.encode("latin-1")
    b2 = seq2 if isinstance(seq2, bytes) else seq2.encode("latin-1")
    out = _abi.AtrMergeResult()
    rc = lib().sim_merge_overlap(b1, len(b1), b2, len(b2), int(bool(insert_matched)), float(min_overlap), float(error_rate),
                                 C.byref(out))
    assert rc == 0
    return out


def locate(read, desc, route=0, lo=0, hi=None, fold_case=False, prev=None, adapter_index=0):
    """Returns (result, used_k1a): result = None | (astart, astop, rstart, rstop, matches, errors) | 'INVALID'."""
    rb = read if isinstance(read, bytes) else read.encode("latin-1")
    out = _abi.AtrMatch()
    reduce = 0
    if prev is not None:
        out = prev
        reduce = 1
    used = C.c_int(0)
    rc = lib().sim_locate(C.byref(desc), adapter_index, reduce, rb, len(rb), lo, len(rb) if hi is None else hi,
                          int(fold_case), route, C.byref(out), C.byref(used))
    if rc != 0:
        raise RuntimeError("sim_locate rc=%d" % rc)
    return decode(out), used.value, out


def decode(m):
    if m.status == _abi.ATR_ST_NONE:
        return None
    if m.status == _abi.ATR_ST_INVALID:
        return "INVALID"
    return (m.astart, m.astop, m.rstart, m.rstop, m.matches, m.errors)


def match_insert(desc, r1, r2, route=0):
    """Returns (record (INSERT_DTYPE scalar), used_packed)."""
    import numpy as np
    b1 = r1 if isinstance(r1, bytes) else r1.encode("latin-1")
    b2 = r2 if isinstance(r2, bytes) else r2.encode("latin-1")
    out = np.zeros(1, dtype=_abi.INSERT_DTYPE)
    used = C.c_int(0)
    rc = lib().sim_match_insert(C.byref(desc), b1, len(b1), b2, len(b2), route, out.ctypes.data, C.byref(used))
    if rc != 0:
        raise RuntimeError("sim_match_insert rc=%d" % rc)
    return out[0], bool(used.value)


def multi_locate(ref, query, rate, flags, min_overlap, max_matches=100):
    r, q = ref.encode("ascii"), query.encode("ascii")
    out = (C.c_int * (6 * (max_matches + len(r) + 2)))()
    cnt = lib().sim_multi_locate(r, len(r), q, len(q), float(rate), flags, min_overlap, max_matches, out)
    if cnt == 0:
        return None
    return [tuple(out[6 * t:6 * t + 6]) for t in range(cnt)]


def trim_fastq(text, adapters, times=1, max_len=512, final=True, **read_ops):
    """CPU run of the FASTQ path's device functions. adapters: atropos_b200.adapters.Adapter objects.
    Returns (out bytes, TrimStats, consumed) or raises atropos_b200.fastq.FormatError."""
    import numpy as np
    from atropos_b200 import fastq
    from atropos_b200.adapters import LINKED
    back_ref = None
    if len(adapters) == 1 and getattr(adapters[0], "where", None) == LINKED:
        linked = adapters[0]
        adapters = [linked.front_adapter, linked.back_adapter]
        back_desc, back_keep = linked.back_adapter.descriptor()
        back_ref = C.byref(back_desc)
        descs, keep = zip(*[linked.front_adapter.descriptor()])
        n_front = 1
    else:
        descs, keep = zip(*[a.descriptor() for a in adapters])
        n_front = len(descs)
    arr = (_abi.AtrAdapterDesc * len(descs))(*descs)
    max_errors = max(int(a.max_error_rate * len(a.sequence)) for a in adapters)
    stats = fastq.TrimStats(len(adapters), max_len, max_errors)
    opts = _abi.AtrTrimOpts(times, max_len, max_errors, int(bool(final)), 0, None, _abi.make_read_ops(**read_ops))
    oc = SimOpsCounters()
    out = np.empty(len(text) + 1, dtype=np.uint8)
    counters = np.zeros(5, dtype=np.int64)
    nout, consumed = C.c_longlong(0), C.c_longlong(0)
    err = _abi.AtrFastqError()
    rc = lib().sim_trim_fastq(arr, n_front, C.byref(opts), text, len(text), out.ctypes.data, C.byref(nout),
                              C.byref(consumed), counters.ctypes.data, stats.errors_front.ctypes.data,
                              stats.errors_back.ctypes.data, stats.adjacent.ctypes.data, C.byref(err), C.byref(oc), back_ref)
    if rc == _abi.ATR_E_FORMAT:
        raise fastq.FormatError(fastq.format_error_message(np.frombuffer(text, dtype=np.uint8), err))
    if rc != 0:
        raise RuntimeError("sim_trim_fastq rc=%d" % rc)
    fastq._add_ops_stats(stats.ops, oc)
    stats.records, stats.with_adapters, stats.bp_in, stats.bp_out, stats.overflow = (int(x) for x in counters)
    return bytes(out[:nout.value]), stats, consumed.value


def trim_fastq_pe(text1, text2, adapter1, adapter2, insert_aligner=None, symmetric=True, min_insert_overlap=1, max_len=256,
                  final=True, times=1, mismatch_action=None, merge_overlapping=False, merge_min_overlap=0.9, merge_error_rate=0.2,
                  merged_output=True, **read_ops):
    """CPU run of the paired-end FASTQ path's device functions (insert mode, or adapter mode when insert_aligner is
    None and adapter1 / adapter2 are lists). Returns ((out1, out2), PairTrimStats, consumed)."""
    import numpy as np
    from atropos_b200 import fastq
    as_list = lambda a: [] if a is None else (list(a) if isinstance(a, (list, tuple)) else [a])
    ads = [as_list(adapter1), as_list(adapter2)]
    keep, arrs = [], []
    for lst in ads:
        pairs = [a.descriptor() for a in lst]
        keep.append(pairs)
        arrs.append((_abi.AtrAdapterDesc * max(1, len(pairs)))(*[d for d, _ in pairs]))
    if insert_aligner is not None:
        idesc, k3 = insert_aligner.descriptor(max_len)
        iref = C.byref(idesc)
        max_errors = max(len(a.sequence) for a in ads[0] + ads[1])
    else:
        iref = None
        max_errors = max([int(a.max_error_rate * len(a.sequence)) for a in ads[0] + ads[1]] or [0])
    stats = fastq.PairTrimStats(max_len, max_errors, (len(ads[0]), len(ads[1])))
    opts = _abi.AtrTrimPeOpts(int(symmetric), min_insert_overlap, max_len, max_errors, int(bool(final)), times,
                              _abi.MISMATCH_ACTIONS[mismatch_action], 0, 0, _abi.make_read_ops(**read_ops))
    corrected = (C.c_longlong * 3)()
    oc = SimOpsCounters()
    o1 = np.empty(len(text1) + 1, dtype=np.uint8)
    o2 = np.empty(len(text2) + 1, dtype=np.uint8)
    counters = np.zeros(9, dtype=np.int64)
    nout, consumed = (C.c_longlong * 3)(), (C.c_longlong * 2)()
    err = _abi.AtrFastqError()
    mo = _abi.AtrMergeOpts(float(merge_min_overlap), float(merge_error_rate)) if merge_overlapping else None
    om = np.empty(len(text1) + len(text2) + 1, dtype=np.uint8)
    mcounters = (C.c_longlong * 6)()
    rc = lib().sim_trim_fastq_pe(iref, arrs[0], len(ads[0]), arrs[1], len(ads[1]), stats.errors_front[0].ctypes.data,
                                 stats.errors_front[1].ctypes.data, C.byref(opts), text1, len(text1), text2, len(text2),
                                 o1.ctypes.data, o2.ctypes.data, nout, consumed, counters.ctypes.data,
                                 stats.errors_back[0].ctypes.data, stats.errors_back[1].ctypes.data,
                                 stats.adjacent[0].ctypes.data, stats.adjacent[1].ctypes.data, C.byref(err), C.byref(oc), corrected,
                                 C.byref(mo) if mo is not None else None, om.ctypes.data if merged_output else None, mcounters)
    if rc == _abi.ATR_E_FORMAT:
        raise fastq.FormatError(fastq.format_error_message((np.frombuffer(text1, dtype=np.uint8),
                                                            np.frombuffer(text2, dtype=np.uint8)), err))
    if rc != 0:
        raise RuntimeError("sim_trim_fastq_pe rc=%d" % rc)
    fastq._add_ops_stats(stats.ops, oc)
    c = [int(x) for x in counters]
    stats.records, stats.insert_matches, stats.overflow = c[0], c[1], c[8]
    stats.with_adapters, stats.bp_in, stats.bp_out = [c[2], c[3]], [c[4], c[5]], [c[6], c[7]]
    stats.records_corrected, stats.bp_corrected = int(corrected[0]), [int(corrected[1]), int(corrected[2])]
    if merge_overlapping:
        stats.merged, stats.merged_written, stats.bp_merged_written = int(mcounters[0]), int(mcounters[1]), int(mcounters[2])
        stats.merge_records_corrected, stats.merge_bp_corrected = int(mcounters[3]), [int(mcounters[4]), int(mcounters[5])]
        return (bytes(o1[:nout[0]]), bytes(o2[:nout[1]]), bytes(om[:nout[2]]) if merged_output else b""), stats, (consumed[0], consumed[1])
    return (bytes(o1[:nout[0]]), bytes(o2[:nout[1]])), stats, (consumed[0], consumed[1])
