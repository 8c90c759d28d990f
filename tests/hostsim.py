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
        _lib = L
    return _lib


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
