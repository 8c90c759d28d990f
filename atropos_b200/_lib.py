"""ctypes binding of libatropos_b200.so (the C ABI in include/atropos_b200.h).

There is no CPU fallback: if the CUDA library has not been built, or no CUDA device is present,
importing/using the engine fails loudly.
"""
import ctypes as C
import os

from . import _abi

PKG = os.path.dirname(os.path.abspath(__file__))
# ATROPOS_B200_LIB: another build of the same library (developer A/B measurements of kernel variants); still the CUDA library
LIB_PATH = os.environ.get("ATROPOS_B200_LIB") or os.path.join(PKG, "libatropos_b200.so")

_lib = None

# every symbol include/atropos_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "atr_abi_version", "atr_device_count", "atr_ctx_create", "atr_ctx_destroy", "atr_last_error", "atr_ctx_sync",
    "atr_ctx_stream", "atr_ctx_launch_count", "atr_ctx_last_kernel_ms", "atr_ctx_set_profiling",
    "atr_ctx_last_phase_ms", "atr_ctx_last_phase_name", "atr_adapterset_create",
    "atr_adapterset_destroy", "atr_packed_words", "atr_pack_device", "atr_pack_reads_host", "atr_locate_batch_device",
    "atr_locate_batch_host", "atr_locate_batch_host_packed", "atr_compare_prefixes", "atr_insertset_create", "atr_insertset_destroy",
    "atr_match_insert_batch_device", "atr_match_insert_batch_host", "atr_multi_locate", "atr_merge_overlap_batch_host", "atr_trim_fastq_host", "atr_trim_fastq_pe_host", "atr_trim_fastq_pe_merge_host",
]


class EngineError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises ImportError if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "atropos_b200: %s is missing. Build it with `python -m atropos_b200.build` (needs nvcc); "
            "this engine has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.atr_abi_version.restype = C.c_int
    L.atr_device_count.restype = C.c_int
    L.atr_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.atr_ctx_destroy.argtypes = [vp]
    L.atr_ctx_destroy.restype = None
    L.atr_last_error.argtypes = [vp]
    L.atr_last_error.restype = C.c_char_p
    L.atr_ctx_sync.argtypes = [vp]
    L.atr_ctx_stream.argtypes = [vp]
    L.atr_ctx_stream.restype = vp
    L.atr_ctx_launch_count.argtypes = [vp, C.c_int]
    L.atr_ctx_launch_count.restype = i64
    L.atr_ctx_last_kernel_ms.argtypes = [vp]
    L.atr_ctx_last_kernel_ms.restype = C.c_float
    L.atr_ctx_set_profiling.argtypes = [vp, C.c_int]
    L.atr_ctx_last_phase_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    L.atr_ctx_last_phase_name.argtypes = [vp, C.c_int]
    L.atr_ctx_last_phase_name.restype = C.c_char_p
    L.atr_adapterset_create.argtypes = [vp, i32, C.POINTER(_abi.AtrAdapterDesc), C.POINTER(vp)]
    L.atr_adapterset_destroy.argtypes = [vp]
    L.atr_adapterset_destroy.restype = None
    L.atr_packed_words.argtypes = [vp, i64]
    L.atr_packed_words.restype = i64
    L.atr_pack_device.argtypes = [vp, vp, vp, i64, C.c_int, vp, vp, vp]
    L.atr_locate_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, i64, vp]
    L.atr_locate_batch_host.argtypes = [vp, vp, vp, vp, vp, i64, C.c_int, vp]
    L.atr_pack_reads_host.argtypes = [vp, vp, i64, C.c_int, C.c_int, vp, vp, vp]
    L.atr_locate_batch_host_packed.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, i64, vp]
    L.atr_compare_prefixes.argtypes = [vp, C.c_char_p, i32, C.c_char_p, i32, C.c_int, C.c_int, C.POINTER(i32)]
    L.atr_insertset_create.argtypes = [vp, C.POINTER(_abi.AtrInsertDesc), C.POINTER(vp)]
    L.atr_insertset_destroy.argtypes = [vp]
    L.atr_insertset_destroy.restype = None
    L.atr_match_insert_batch_device.argtypes = [vp] * 12 + [i64, vp]
    L.atr_match_insert_batch_host.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp]
    L.atr_multi_locate.argtypes = [vp, C.c_char_p, i32, C.c_char_p, i32, f64, i32, i32, i32, C.POINTER(i32),
                                   C.POINTER(i32)]
    L.atr_merge_overlap_batch_host.argtypes = [vp, vp, vp, vp, vp, vp, i64, f64, f64, vp]
    L.atr_trim_fastq_host.argtypes = [vp, vp, C.POINTER(_abi.AtrTrimOpts), vp, i64, vp, i64, C.POINTER(i64), C.POINTER(i64),
                                      C.POINTER(_abi.AtrTrimStats), C.POINTER(_abi.AtrFastqError)]
    L.atr_trim_fastq_pe_host.argtypes = [vp, vp, vp, vp, C.POINTER(_abi.AtrTrimPeOpts), vp, i64, vp, i64, vp, i64, vp, i64,
                                         C.POINTER(i64), C.POINTER(i64), C.POINTER(_abi.AtrTrimPeStats),
                                         C.POINTER(_abi.AtrFastqError)]
    L.atr_trim_fastq_pe_merge_host.argtypes = [vp, vp, vp, vp, C.POINTER(_abi.AtrTrimPeOpts), C.POINTER(_abi.AtrMergeOpts), vp, i64, vp, i64,
                                               vp, i64, vp, i64, vp, i64, C.POINTER(i64), C.POINTER(i64),
                                               C.POINTER(_abi.AtrTrimPeStats), C.POINTER(_abi.AtrMergeStats),
                                               C.POINTER(_abi.AtrFastqError)]
    if L.atr_abi_version() != _abi.ATR_ABI_VERSION:
        raise ImportError("atropos_b200: ABI version mismatch, rebuild with `python -m atropos_b200.build --force`")
    _lib = L
    return L


def check(rc, ctx_handle=None):
    """Translate a negative return code into the exception the reference would raise."""
    if rc == 0:
        return
    L = load()
    msg = L.atr_last_error(ctx_handle)
    msg = msg.decode("utf-8", "replace") if msg else "error %d" % rc
    if rc == _abi.ATR_E_ARG:
        raise ValueError(msg)
    if rc == _abi.ATR_E_NOMEM:
        raise MemoryError(msg)
    if rc == _abi.ATR_E_LIMIT:
        raise OverflowError(msg)
    raise EngineError(msg)
