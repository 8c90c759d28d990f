"""TEST-ONLY stand-in for the GPU engine: the product's `__host__ __device__` per-read functions compiled for the CPU
(tests/host_sim, see tests/hostsim.py) behind the method names of atropos_b200.engine.

Why it exists: this container has no GPU, and the reference-side binding (atropos_b200/integration.py) is Python glue
whose correctness -- staging of the modifier chain, replay of precomputed records, windows between rounds -- does not
depend on where the per-read functions run. `install()` lets the `-m "not gpu"` suite run the reference's own tests
through that glue here; on the B200 box tests/test_gpu_reference_suite.py runs the same tests on the real engine.
Never imported by atropos_b200 (tests/test_abi.py::test_product_never_imports_oracle also checks for this name).
"""
import ctypes as C

import numpy as np

import hostsim
from atropos_b200 import _abi, engine


class SimContext(object):
    handle = "sim"
    device = 0

    def compare_prefixes(self, ref, query, wildcard_ref=False, wildcard_query=False):
        r, q = ref.encode("ascii"), query.encode("ascii")
        out = (C.c_int * 6)()
        assert hostsim.lib().sim_compare_prefixes(r, len(r), q, len(q), int(bool(wildcard_ref)), int(bool(wildcard_query)),
                                                  out) == 0
        return tuple(out)

    def multi_locate(self, reference, query, max_error_rate, flags, min_overlap, max_matches=100):
        return hostsim.multi_locate(reference, query, max_error_rate, flags, min_overlap, max_matches)

    def merge_overlap_host(self, ascii1, offsets1, ascii2, offsets2, min_overlap, error_rate, insert_matched=None, out=None):
        n = len(offsets1) - 1
        if out is None:
            out = np.empty(n, dtype=_abi.MERGE_DTYPE)
        b1 = np.ascontiguousarray(ascii1, dtype=np.uint8).tobytes()
        b2 = np.ascontiguousarray(ascii2, dtype=np.uint8).tobytes()
        for i in range(n):
            r = hostsim.merge_overlap(b1[offsets1[i]:offsets1[i + 1]], b2[offsets2[i]:offsets2[i + 1]],
                                      bool(insert_matched[i]) if insert_matched is not None else False, min_overlap, error_rate)
            out[i] = (r.r2_start, r.r2_stop, r.r1_start, r.r1_stop, r.matches, r.errors, r.min_overlap, r.status, r.action)
        return out


_CTX = SimContext()


class SimAdapterSet(object):
    def __init__(self, ctx, descs_and_keep):
        self.ctx = ctx
        self._descs = [d for d, _ in descs_and_keep]
        self._keep = [k for _, k in descs_and_keep]
        self.n_adapters = len(self._descs)
        self.handle = "sim"
        for d in self._descs:                                  # argument checks of atr_adapterset_create
            if d.min_overlap < 1 or d.indel_cost < 1 or d.length < 1:
                raise ValueError("bad adapter descriptor")

    def close(self):
        pass

    def locate_host(self, ascii, offsets, win=None, fold_case=False, out=None):
        ascii = np.ascontiguousarray(ascii, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        if out is None:
            out = np.empty(n, dtype=_abi.MATCH_DTYPE)
        blob = ascii.tobytes()
        L = hostsim.lib()
        for i in range(n):
            read = blob[offsets[i]:offsets[i + 1]]
            lo, hi = (0, len(read)) if win is None else (int(win[i][0]), int(win[i][1]))
            m = _abi.AtrMatch()
            used = C.c_int(0)
            for a, d in enumerate(self._descs):
                rc = L.sim_locate(C.byref(d), a, int(a > 0), read, len(read), lo, hi, int(bool(fold_case)), 0, C.byref(m),
                                  C.byref(used))
                if rc != 0:
                    raise RuntimeError("sim_locate rc=%d" % rc)
            out[i] = (m.astart, m.astop, m.rstart, m.rstop, m.matches, m.errors, m.adapter, m.status)
        return out


class SimInsertSet(object):
    def __init__(self, ctx, desc, keep):
        self.ctx, self._desc, self._keep, self.max_len, self.handle = ctx, desc, keep, desc.max_len, "sim"

    def close(self):
        pass

    def match_insert_host(self, ascii1, offsets1, ascii2, offsets2, out=None):
        n = len(offsets1) - 1
        if out is None:
            out = np.empty(n, dtype=_abi.INSERT_DTYPE)
        b1 = np.ascontiguousarray(ascii1, dtype=np.uint8).tobytes()
        b2 = np.ascontiguousarray(ascii2, dtype=np.uint8).tobytes()
        for i in range(n):
            rec, _ = hostsim.match_insert(self._desc, b1[offsets1[i]:offsets1[i + 1]], b2[offsets2[i]:offsets2[i + 1]])
            out[i] = rec
        return out


def install():
    """Route atropos_b200.engine's three device-backed classes to the CPU simulation (this process only)."""
    engine.default_context = lambda device=0: _CTX
    engine.AdapterSet = SimAdapterSet
    engine.InsertSet = SimInsertSet
    return True
