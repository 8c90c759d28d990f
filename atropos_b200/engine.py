"""Thin object layer over the C ABI: Context (one per GPU), AdapterSet, InsertSet and the batch calls.

Batches are (ascii, offsets) pairs: `ascii` = uint8 array with all reads back to back, `offsets` =
int64[n+1]. `encode_reads` builds them from a list of str. Results come back as numpy structured
arrays (`_abi.MATCH_DTYPE` / `_abi.INSERT_DTYPE`) -- one record per read, no Python objects.
"""
import ctypes as C
import os
import threading

import numpy as np

from . import _abi, _lib

_default_ctx = {}
_lock = threading.Lock()


def encode_reads(reads):
    """list of str/bytes -> (uint8 array, int64 offsets[n+1]). Raises UnicodeEncodeError on non-ASCII
    like the reference (`query.encode('ascii')`, _align.pyx:281)."""
    n = len(reads)
    offsets = np.zeros(n + 1, dtype=np.int64)
    if n == 0:
        return np.zeros(0, dtype=np.uint8), offsets
    if isinstance(reads[0], bytes):
        blob = b"".join(reads)
    else:
        blob = "".join(reads).encode("ascii")
    np.cumsum(np.fromiter((len(r) for r in reads), dtype=np.int64, count=n), out=offsets[1:])
    return np.frombuffer(blob, dtype=np.uint8), offsets


def pack_reads_host(ascii, offsets, fold_case=False, threads=0, out=None):
    """atr_pack_reads_host: (ascii, offsets) -> (codes uint32[], woff uint32[n+1], len uint16[n]) in host memory, the
    layout of include/atropos_b200.h (no GPU involved). `out`: optional (codes, woff, len) arrays to fill, e.g. pinned."""
    L = _lib.load()
    ascii = np.ascontiguousarray(ascii, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    if out is None:
        words = int(L.atr_packed_words(offsets.ctypes.data, n))
        out = (np.empty(words + 8, dtype=np.uint32), np.empty(n + 1, dtype=np.uint32), np.empty(n, dtype=np.uint16))
    codes, woff, lens = out
    rc = L.atr_pack_reads_host(ascii.ctypes.data if ascii.size else None, offsets.ctypes.data, n, int(bool(fold_case)),
                               int(threads), codes.ctypes.data, woff.ctypes.data, lens.ctypes.data)
    if rc != 0:
        raise ValueError("atr_pack_reads_host failed (%d): a read longer than 32767 nt or bad offsets" % rc)
    return codes, woff, lens


def fixed_length_offsets(n, length):
    return np.arange(n + 1, dtype=np.int64) * int(length)


class Context(object):
    """One per GPU per host thread (atr_ctx)."""

    def __init__(self, device=0):
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.atr_ctx_create(int(device), C.byref(h)), None)
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self._L.atr_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self._L.atr_ctx_sync(self.handle), self.handle)

    @property
    def stream(self):
        return self._L.atr_ctx_stream(self.handle)

    def launch_count(self, reset=False):
        return int(self._L.atr_ctx_launch_count(self.handle, int(reset)))

    def last_kernel_ms(self):
        return float(self._L.atr_ctx_last_kernel_ms(self.handle))

    def set_profiling(self, on=True):
        _lib.check(self._L.atr_ctx_set_profiling(self.handle, int(bool(on))), self.handle)

    def last_phase_ms(self):
        """[filter, refine, band, wide] kernel times (ms) of the last device call on the fast path, or []"""
        buf = (C.c_float * 4)()
        k = self._L.atr_ctx_last_phase_ms(self.handle, buf, 4)
        return [float(buf[i]) for i in range(k)]

    def last_phase_names(self):
        """the kernels behind the four intervals of last_phase_ms()"""
        return [(self._L.atr_ctx_last_phase_name(self.handle, i) or b"").decode() for i in range(4)]

    # ---- single-call functions ----
    def compare_prefixes(self, ref, query, wildcard_ref=False, wildcard_query=False):
        r, q = ref.encode("ascii"), query.encode("ascii")
        out = (C.c_int32 * 6)()
        _lib.check(self._L.atr_compare_prefixes(self.handle, r, len(r), q, len(q), int(bool(wildcard_ref)),
                                                int(bool(wildcard_query)), out), self.handle)
        return tuple(out)

    def multi_locate(self, reference, query, max_error_rate, flags, min_overlap, max_matches=100):
        r, q = reference.encode("ascii"), query.encode("ascii")
        out = (C.c_int32 * (6 * (max_matches + len(r) + 2)))()
        cnt = C.c_int32(0)
        _lib.check(self._L.atr_multi_locate(self.handle, r, len(r), q, len(q), float(max_error_rate), int(flags),
                                            int(min_overlap), int(max_matches), out, C.byref(cnt)), self.handle)
        if cnt.value == 0:
            return None
        return [tuple(out[6 * t:6 * t + 6]) for t in range(cnt.value)]

    def merge_overlap_host(self, ascii1, offsets1, ascii2, offsets2, min_overlap, error_rate, insert_matched=None, out=None):
        """atr_merge_overlap_batch_host: one MERGE_DTYPE record per pair."""
        ascii1 = np.ascontiguousarray(ascii1, dtype=np.uint8)
        ascii2 = np.ascontiguousarray(ascii2, dtype=np.uint8)
        offsets1 = np.ascontiguousarray(offsets1, dtype=np.int64)
        offsets2 = np.ascontiguousarray(offsets2, dtype=np.int64)
        n = len(offsets1) - 1
        assert len(offsets2) - 1 == n
        if out is None:
            out = np.empty(n, dtype=_abi.MERGE_DTYPE)
        im = None
        if insert_matched is not None:
            im = np.ascontiguousarray(insert_matched, dtype=np.uint8)
            assert im.shape == (n,)
        _lib.check(self._L.atr_merge_overlap_batch_host(
            self.handle, ascii1.ctypes.data if ascii1.size else None, offsets1.ctypes.data,
            ascii2.ctypes.data if ascii2.size else None, offsets2.ctypes.data, im.ctypes.data if im is not None else None,
            n, float(min_overlap), float(error_rate), out.ctypes.data), self.handle)
        return out


def context_key(device=0):
    """An atr_ctx owns staging buffers and is not re-entrant (include/atropos_b200.h: one per GPU per host
    thread), and a CUDA context does not survive fork(): the implicit contexts are cached per
    (device, process, thread)."""
    return (int(device), os.getpid(), threading.get_ident())


def default_context(device=0):
    key = context_key(device)
    with _lock:
        ctx = _default_ctx.get(key)
        if ctx is None or ctx.handle is None:
            for k in [k for k in _default_ctx if k[1] != key[1]]:
                # inherited from the parent of a fork(): the handle means nothing here and must not be destroyed
                _default_ctx.pop(k).handle = None
            ctx = _default_ctx[key] = Context(device)
        return ctx


class AdapterSet(object):
    """atr_adapterset: the aligner state of one or several adapters (AdapterCutter.adapters order)."""

    def __init__(self, ctx, descs_and_keep):
        self.ctx = ctx
        self._L = ctx._L
        self._keep = [k for _, k in descs_and_keep]
        arr = (_abi.AtrAdapterDesc * len(descs_and_keep))(*[d for d, _ in descs_and_keep])
        h = C.c_void_p()
        _lib.check(self._L.atr_adapterset_create(ctx.handle, len(descs_and_keep), arr, C.byref(h)), ctx.handle)
        self.handle = h
        self.n_adapters = len(descs_and_keep)

    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self._L.atr_adapterset_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def locate_host(self, ascii, offsets, win=None, fold_case=False, out=None):
        """atr_locate_batch_host: returns a MATCH_DTYPE array of len(offsets)-1 records."""
        ascii = np.ascontiguousarray(ascii, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        if out is None:
            out = np.empty(n, dtype=_abi.MATCH_DTYPE)
        wptr = None
        if win is not None:
            win = np.ascontiguousarray(win, dtype=np.uint16)
            assert win.shape == (n, 2)
            wptr = win.ctypes.data
        _lib.check(self._L.atr_locate_batch_host(self.ctx.handle, self.handle, ascii.ctypes.data if ascii.size else None,
                                                 offsets.ctypes.data, wptr, n, int(bool(fold_case)), out.ctypes.data),
                   self.ctx.handle)
        return out

    def locate_host_packed(self, codes, woff, lens, win=None, ascii=None, offsets=None, fold_case=False, out=None):
        """atr_locate_batch_host_packed: reads packed in host memory (pack_reads_host); ascii / offsets only serve the
        escaped reads."""
        n = len(lens)
        if out is None:
            out = np.empty(n, dtype=_abi.MATCH_DTYPE)
        wptr = None
        if win is not None:
            win = np.ascontiguousarray(win, dtype=np.uint16)
            assert win.shape == (n, 2)
            wptr = win.ctypes.data
        a_ptr = o_ptr = None
        if ascii is not None and offsets is not None:
            ascii = np.ascontiguousarray(ascii, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            a_ptr, o_ptr = (ascii.ctypes.data if ascii.size else None), offsets.ctypes.data
        _lib.check(self._L.atr_locate_batch_host_packed(self.ctx.handle, self.handle, codes.ctypes.data, woff.ctypes.data,
                                                        lens.ctypes.data, wptr, a_ptr, o_ptr, int(bool(fold_case)), int(n),
                                                        out.ctypes.data), self.ctx.handle)
        return out

    def locate_device(self, d_codes, d_woff, d_len, n, d_out, d_win=None, d_ascii=None, d_offsets=None,
                      fold_case=False):
        """atr_locate_batch_device with raw device pointers (ints). Asynchronous on ctx.stream."""
        _lib.check(self._L.atr_locate_batch_device(self.ctx.handle, self.handle, d_codes, d_woff, d_len, d_win, d_ascii,
                                                   d_offsets, int(bool(fold_case)), int(n), d_out), self.ctx.handle)


class InsertSet(object):
    """atr_insertset: the state of one InsertAligner."""

    def __init__(self, ctx, desc, keep):
        self.ctx = ctx
        self._L = ctx._L
        self._keep = keep
        h = C.c_void_p()
        _lib.check(self._L.atr_insertset_create(ctx.handle, C.byref(desc), C.byref(h)), ctx.handle)
        self.handle = h
        self.max_len = desc.max_len

    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self._L.atr_insertset_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def match_insert_host(self, ascii1, offsets1, ascii2, offsets2, out=None):
        ascii1 = np.ascontiguousarray(ascii1, dtype=np.uint8)
        ascii2 = np.ascontiguousarray(ascii2, dtype=np.uint8)
        offsets1 = np.ascontiguousarray(offsets1, dtype=np.int64)
        offsets2 = np.ascontiguousarray(offsets2, dtype=np.int64)
        n = len(offsets1) - 1
        assert len(offsets2) - 1 == n
        if out is None:
            out = np.empty(n, dtype=_abi.INSERT_DTYPE)
        _lib.check(self._L.atr_match_insert_batch_host(
            self.ctx.handle, self.handle, ascii1.ctypes.data if ascii1.size else None, offsets1.ctypes.data,
            ascii2.ctypes.data if ascii2.size else None, offsets2.ctypes.data, n, out.ctypes.data), self.ctx.handle)
        return out

    def match_insert_device(self, c1, w1, l1, c2, w2, l2, n, d_out, a1=None, o1=None, a2=None, o2=None):
        _lib.check(self._L.atr_match_insert_batch_device(self.ctx.handle, self.handle, c1, w1, l1, c2, w2, l2,
                                                         a1, o1, a2, o2, int(n), d_out), self.ctx.handle)
