"""The reference-side binding as real code: run the UNMODIFIED jdidion/atropos package on top of this engine.

Two levels, both installed by :func:`install` BEFORE ``atropos.align`` is first imported:

1. **per-call binding** -- a meta-path finder serves the reference's compiled extension module
   ``atropos.align._align`` (``atropos/align/__init__.py:6``: ``from atropos.align._align import Aligner,
   MultiAligner, compare_prefixes, locate``) from :mod:`atropos_b200.align`.  Every ``Aligner.locate`` /
   ``MultiAligner.locate`` / ``compare_prefixes`` of the reference (``adapters/__init__.py:311-322, 370-382``,
   ``align/__init__.py:230-233, 285-288, 351``) then is one call through the C ABI into the CUDA kernels.
   Correct, and slow (one launch + sync per read): it exists so that everything the batched binding does not
   take (colorspace adapters, exotic modifier orders) still runs on the GPU, bit-exactly.

2. **batched binding** -- ``TrimPipeline.handle_records`` (``commands/trim/__init__.py:82-84`` over
   ``commands/base.py:65-78``) is replaced by a staged version: the modifiers in front of the adapter stage run
   over the whole batch, ONE set of batch calls computes every adapter match of the batch
   (``AdapterCutter``: ``atr_locate_batch_host`` per round, all adapters of the cutter as one panel;
   ``InsertAdapterCutter``: ``atr_match_insert_batch_host`` + the two fallback alignments where needed), then the
   reference's own ``AdapterCutter.__call__`` / ``InsertAdapterCutter.__call__`` run per record exactly as
   written, with ``_best_match`` / ``aligner.match_insert`` / ``adapter.match_to`` *replaying* the precomputed
   records as the reference's own ``Match`` objects.  Trimming, masking, error correction, statistics, filters and
   formatters are the reference's code and see identical ``Match`` objects.

Nothing here computes an alignment on the CPU: without the CUDA library or a device every path raises
(:class:`atropos_b200._lib.EngineError`).
"""
import importlib.abc
import importlib.machinery
import sys

import numpy as np

from . import _abi, engine

SHIM_NAME = "atropos.align._align"

#: counters a caller (and tests/test_gpu_reference_suite.py) can read to see which binding did the work
STATS = {"percall_locate": 0, "percall_multi_locate": 0, "percall_compare_prefixes": 0,
         "batched_batches": 0, "batched_reads": 0, "batched_gpu_calls": 0, "percall_batches": 0,
         "batched_merge_batches": 0}

_installed = {"finder": False, "batched": False}
_orig_handle_records = None


# ------------------------------------------------------------------------------------------------
# 1. per-call binding: atropos.align._align served by atropos_b200.align
# ------------------------------------------------------------------------------------------------
class _AlignShimFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """sys.meta_path entry: ``import atropos.align._align`` -> the GPU-backed classes."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != SHIM_NAME:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, origin="atropos_b200.align (CUDA, sm_100a)")

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        from . import align as gpu

        class Aligner(gpu.Aligner):
            __doc__ = gpu.Aligner.__doc__

            def locate(self, query):
                STATS["percall_locate"] += 1
                return gpu.Aligner.locate(self, query)

        class MultiAligner(gpu.MultiAligner):
            __doc__ = gpu.MultiAligner.__doc__

            def locate(self, reference, query, max_matches=100):
                STATS["percall_multi_locate"] += 1
                return gpu.MultiAligner.locate(self, reference, query, max_matches)

        def compare_prefixes(ref, query, wildcard_ref=False, wildcard_query=False):
            STATS["percall_compare_prefixes"] += 1
            return gpu.compare_prefixes(ref, query, wildcard_ref, wildcard_query)

        def locate(reference, query, max_error_rate, flags=gpu.SEMIGLOBAL, wildcard_ref=False, wildcard_query=False,
                   min_overlap=1):
            aligner = Aligner(reference, max_error_rate, flags, wildcard_ref, wildcard_query)
            aligner.min_overlap = min_overlap
            return aligner.locate(query)

        module.Aligner = Aligner
        module.MultiAligner = MultiAligner
        module.compare_prefixes = compare_prefixes
        module.locate = locate
        module.__atropos_b200__ = True


_FINDER = _AlignShimFinder()


def shim_active():
    """True iff ``atropos.align._align`` is (or will be) the GPU module of this package."""
    mod = sys.modules.get(SHIM_NAME)
    if mod is not None:
        return bool(getattr(mod, "__atropos_b200__", False))
    return _installed["finder"]


def install(batched=True):
    """Install the per-call binding and, if `batched`, the batched trim pipeline. Must run before anything imports
    ``atropos.align`` (the reference binds the four names at import time)."""
    mod = sys.modules.get(SHIM_NAME)
    if mod is not None and not getattr(mod, "__atropos_b200__", False):
        raise RuntimeError("atropos.align._align is already imported from %r: call atropos_b200.integration.install() "
                           "before importing atropos" % (getattr(mod, "__file__", None),))
    if not _installed["finder"]:
        sys.meta_path.insert(0, _FINDER)
        _installed["finder"] = True
    if batched:
        install_batched_pipeline()


def uninstall():
    """Remove the finder and restore ``TrimPipeline.handle_records`` (already imported modules keep what they bound)."""
    global _orig_handle_records
    if _installed["finder"]:
        sys.meta_path[:] = [f for f in sys.meta_path if f is not _FINDER]
        _installed["finder"] = False
    if _installed["batched"]:
        import atropos.commands.trim as trim
        trim.TrimPipeline.handle_records = _orig_handle_records
        _installed["batched"] = False


# ------------------------------------------------------------------------------------------------
# 2. batched binding
# ------------------------------------------------------------------------------------------------
_REF = None


def _ref():
    """The reference modules the batched binding talks to (imported lazily, through the finder; looked up once: this is
    called per record)."""
    global _REF
    if _REF is None:
        import atropos
        import atropos.adapters as adapters
        import atropos.align as align
        import atropos.commands.base as base
        import atropos.commands.trim as trim
        import atropos.commands.trim.modifiers as modifiers
        _REF = (atropos, adapters, align, base, trim, modifiers)
    return _REF


def adapter_descriptor(adapter):
    """(AtrAdapterDesc, keepalive) of a reference ``Adapter`` (adapters/__init__.py:231-322) with
    ``Adapter.match_to`` semantics; the indel cost lives in its (shim) aligner, like in the reference."""
    m = len(adapter.sequence)
    rmp_ok = None
    if adapter.max_rmp is not None:
        rmp_ok = np.zeros((m + 1, m + 1), dtype=np.uint8)
        for size in range(0, m + 1):
            for matches in range(0, size + 1):
                rmp_ok[size, matches] = adapter.match_probability(matches, size) <= adapter.max_rmp
    return _abi.make_adapter_desc(
        adapter.sequence, adapter.max_error_rate, adapter.where, adapter.adapter_wildcards, adapter.read_wildcards,
        adapter.min_overlap, adapter.aligner._indel_cost, match_to_semantics=True, no_indels=not adapter.indels,
        rmp_ok=rmp_ok)


def _plain(adapter, adapters_mod):
    """exactly the reference's Adapter (not ColorspaceAdapter or another subclass) on top of the shim aligner"""
    return type(adapter) is adapters_mod.Adapter and hasattr(adapter.aligner, "_indel_cost")


class _Desync(RuntimeError):
    pass


class _CutterReplay(object):
    """Precomputed rounds of one AdapterCutter over the batch; stands in for ``_best_match`` (modifiers.py:107-122)."""

    def __init__(self, cutter, linked):
        self.cutter, self.linked = cutter, linked
        self.rounds, self.lens, self.idx, self.t, self.lo, self.hi = [], None, -1, 0, 0, 0

    def seek(self, idx):
        self.idx, self.t = idx, 0
        self.lo, self.hi = 0, self.lens[idx]

    def __call__(self, read):
        _, adapters_mod, align, _, _, _ = _ref()
        i, t = self.idx, self.t
        self.t += 1
        if len(read.sequence) != self.hi - self.lo:
            raise _Desync("batched binding out of step at read %d round %d: the cutter sees %d nt, the batch call "
                          "aligned %d" % (i, t, len(read.sequence), self.hi - self.lo))
        if self.linked is not None:
            return self._linked(read, i)
        if t >= len(self.rounds):
            return None
        # one round = a list of plain tuples in MATCH_DTYPE's field order (numpy record scalars cost ~1 us per field)
        astart, astop, rstart, rstop, matches, errors, a_idx, st = self.rounds[t][i]
        if st == _abi.ATR_ST_NONE:
            return None
        if st == _abi.ATR_ST_INVALID:
            raise ValueError('A Match requires at least one matching position.')
        adapter = self.cutter.adapters[a_idx]
        match = align.Match(astart, astop, rstart, rstop, matches, errors, adapter._front_flag, adapter, read)
        if match.front:                      # what adapter.trimmed(match) leaves for the next round
            self.lo += match.rstop
        else:
            self.hi = self.lo + match.rstart
        return match

    def _linked(self, read, i):
        _, adapters_mod, align, _, _, _ = _ref()
        front, back = self.rounds
        la = self.linked

        def mk(rec, adapter, rd):
            st = int(rec["status"])
            if st == _abi.ATR_ST_NONE:
                return None
            if st == _abi.ATR_ST_INVALID:
                raise ValueError('A Match requires at least one matching position.')
            return align.Match(int(rec["astart"]), int(rec["astop"]), int(rec["rstart"]), int(rec["rstop"]),
                               int(rec["matches"]), int(rec["errors"]), adapter._front_flag, adapter, rd)
        fm = mk(front[i], la.front_adapter, read)
        if fm is None:
            return None
        return adapters_mod.LinkedMatch(fm, mk(back[i], la.back_adapter, read[fm.rstop:]), la)


class _InsertReplay(object):
    """Precomputed results of one InsertAdapterCutter over the batch: stands in for ``aligner.match_insert`` and the
    two ``adapter.match_to`` fallbacks (modifiers.py:391-406)."""

    def __init__(self, cutter):
        self.cutter = cutter
        self.ins = self.fb = None
        self.idx = -1

    def seek(self, idx):
        self.idx = idx

    def match_insert(self, seq1, seq2):
        from .align import InsertAligner
        return InsertAligner.result_from_record(self.ins[self.idx])

    def match_to(self, side):
        _, _, align, _, _, _ = _ref()
        adapter = (self.cutter.adapter1, self.cutter.adapter2)[side]

        def replay(read):
            rec = self.fb[side][self.idx]
            st = int(rec["status"])
            if st == _abi.ATR_ST_NONE:
                return None
            if st == _abi.ATR_ST_INVALID:
                raise ValueError('A Match requires at least one matching position.')
            return align.Match(int(rec["astart"]), int(rec["astop"]), int(rec["rstart"]), int(rec["rstop"]),
                               int(rec["matches"]), int(rec["errors"]), adapter._front_flag, adapter, read)
        return replay


class _MergeReplay(object):
    """Precomputed alignments of one MergeOverlapping modifier over the batch: stands in for the per-pair
    ``Aligner(read2_rc, error_rate, flags).locate(read1.sequence)`` (modifiers.py:886-895); everything else of the
    modifier -- the length test, reverse_complement, error correction, the four ways to merge -- stays reference code."""

    def __init__(self, mod):
        self.mod = mod
        self.recs, self.lens, self.idx = None, None, -1

    def seek(self, idx):
        self.idx = idx

    def precompute(self, mids, device):
        from .modifiers import MergeOverlapping as GpuMerge
        gm = self.mod.__dict__.setdefault("_atr_gpu_merge", {})
        key = engine.context_key(device)
        if key not in gm:
            gm[key] = GpuMerge(self.mod.min_overlap, self.mod.error_rate, None, device=device)
        im = np.fromiter((bool(a.insert_overlap and b.insert_overlap) for a, b in mids), dtype=np.uint8, count=len(mids))
        a1, o1 = engine.encode_reads([a.sequence for a, _ in mids])
        a2, o2 = engine.encode_reads([b.sequence for _, b in mids])
        self.lens = (np.diff(o1), np.diff(o2))
        self.recs = gm[key].align_batch((a1, o1), (a2, o2), insert_matched=im)
        STATS["batched_gpu_calls"] += 1

    def Aligner(self, reference, max_error_rate, flags=15, *args, **kwargs):
        """what MergeOverlapping.__call__ constructs per pair: an object whose locate() gives this pair's alignment"""
        rp = self

        class _Located(object):
            def locate(self, query):
                i = rp.idx
                if len(query) != int(rp.lens[0][i]) or len(reference) != int(rp.lens[1][i]):
                    raise _Desync("batched binding out of step at pair %d: MergeOverlapping sees %d / %d nt, the batch call "
                                  "aligned %d / %d" % (i, len(query), len(reference), int(rp.lens[0][i]), int(rp.lens[1][i])))
                rec = rp.recs[i]
                if int(rec["r2_stop"]) == 0 and int(rec["r1_stop"]) == 0:      # locate() returned None
                    return None
                return tuple(int(rec[f]) for f in ("r2_start", "r2_stop", "r1_start", "r1_stop", "matches", "errors"))
        return _Located()


class _Plan(object):
    """Where the adapter stage sits in a Modifiers chain and how to batch it. None of this changes what the chain
    computes: every modifier sees the records in the reference's order, once."""

    def __init__(self, mods):
        _, adapters_mod, _, _, _, modifiers = _ref()
        self.mods = mods
        self.paired = isinstance(mods, modifiers.PairedEndModifiers)
        self.index = None
        self.kind = None
        self._step_cache = {}
        chain = mods.modifiers
        self.n_chain = len(chain)
        stages = [i for i, m in enumerate(chain)
                  if isinstance(m, modifiers.InsertAdapterCutter) or
                  (isinstance(m, list) and any(isinstance(x, modifiers.AdapterCutter) for x in m))]
        if len(stages) != 1:
            return
        i = stages[0]
        stage = chain[i]
        if isinstance(stage, modifiers.InsertAdapterCutter):
            if type(stage) is not modifiers.InsertAdapterCutter or stage.adapter1 is stage.adapter2:
                return
            if not (_plain(stage.adapter1, adapters_mod) and _plain(stage.adapter2, adapters_mod)):
                return
            self.kind, self.index = "insert", i
            self.replays = [_InsertReplay(stage)]
        else:
            replays = [None, None]
            for side in (0, 1):
                c = stage[side]
                if c is None or type(c) is not modifiers.AdapterCutter or not c.adapters:
                    continue
                linked = None
                if len(c.adapters) == 1 and type(c.adapters[0]) is adapters_mod.LinkedAdapter:
                    la = c.adapters[0]
                    if c.times != 1 or not (_plain(la.front_adapter, adapters_mod) and _plain(la.back_adapter, adapters_mod)):
                        continue
                    linked = la
                elif not all(_plain(a, adapters_mod) for a in c.adapters):
                    continue
                replays[side] = _CutterReplay(c, linked)
            if not any(replays):
                return
            self.kind, self.index = "adapter", i
            self.replays = replays
        # MergeOverlapping behind the adapter stage (always the last modifier, commands/trim/__init__.py:546-552): its
        # per-pair alignment is a second batched GPU stage
        self.merge_index, self.merge = None, None
        for j in range(self.index + 1, len(chain)):
            if type(chain[j]) is modifiers.MergeOverlapping and self.paired:
                self.merge_index, self.merge = j, _MergeReplay(chain[j])
                break

    # -- the chain, split at the adapter stage (SingleEndModifiers.modify / PairedEndModifiers.modify,
    #    modifiers.py:1048-1051, 1096-1105) ----------------------------------------------------------------
    def _steps(self, lo, hi):
        """chain[lo:hi] as a flat list of (kind, callable): 0 = pair modifier, 1 = read 1, 2 = read 2. Built once per
        plan (this runs per record; the generic loop spent its time on isinstance tests and slices)."""
        key = (lo, hi)
        steps = self._step_cache.get(key)
        if steps is None:
            _, _, _, _, _, modifiers = _ref()
            steps = []
            for mods in self.mods.modifiers[lo:hi]:
                if isinstance(mods, modifiers.ReadPairModifier):
                    steps.append((0, mods))
                else:
                    if mods[0] is not None:
                        steps.append((1, mods[0]))
                    if self.paired and mods[1] is not None:
                        steps.append((2, mods[1]))
            self._step_cache[key] = steps
        return steps

    def _run(self, lo, hi, read1, read2):
        for kind, fn in self._steps(lo, hi):
            if kind == 1:
                read1 = fn(read1)
            elif kind == 2:
                read2 = fn(read2)
            else:
                read1, read2 = fn(read1, read2)
        return read1, read2

    def before(self, read1, read2):
        return self._run(0, self.index, read1, read2)

    def from_stage(self, read1, read2):
        read1, read2 = self._run(self.index, self.n_chain, read1, read2)
        return (read1, read2) if self.paired else (read1,)

    def stage_to_merge(self, read1, read2):
        return self._run(self.index, self.merge_index, read1, read2)

    def from_merge(self, read1, read2):
        """MergeOverlapping.__call__ (reference code) with its Aligner swapped for the replay, then whatever follows"""
        _, _, _, _, _, modifiers = _ref()
        saved = modifiers.Aligner
        modifiers.Aligner = self.merge.Aligner
        try:
            return self._run(self.merge_index, self.n_chain, read1, read2)
        finally:
            modifiers.Aligner = saved

    # -- the GPU stage -----------------------------------------------------------------------------------
    def precompute(self, state, device=0):
        stage = self.mods.modifiers[self.index]
        n = len(state)
        if self.kind == "insert":
            self._precompute_insert(stage, state, device)
        else:
            for side in (0, 1):
                rp = self.replays[side]
                if rp is not None:
                    self._precompute_cutter(rp, [s[side].sequence for s in state], device)
        STATS["batched_batches"] += 1
        STATS["batched_reads"] += n

    def _set_of(self, owner, adapters, device):
        sets = owner.__dict__.setdefault("_atr_sets", {})
        key = (engine.context_key(device), tuple(id(a) for a in adapters))
        if key not in sets:
            sets[key] = engine.AdapterSet(engine.default_context(device), [adapter_descriptor(a) for a in adapters])
        return sets[key]

    def _precompute_cutter(self, rp, seqs, device):
        c = rp.cutter
        ascii, offsets = engine.encode_reads(seqs)
        n = len(seqs)
        lens = np.diff(offsets).astype(np.int64)
        rp.lens = lens.tolist()
        if rp.linked is not None:
            la = rp.linked
            front = self._set_of(c, [la.front_adapter], device).locate_host(ascii, offsets, fold_case=True)
            hit = front["status"] == _abi.ATR_ST_MATCH
            win = np.zeros((n, 2), dtype=np.uint16)
            win[:, 0] = np.where(hit, front["rstop"], lens).astype(np.uint16)
            win[:, 1] = lens.astype(np.uint16)
            back = self._set_of(c, [la.back_adapter], device).locate_host(ascii, offsets, win=win, fold_case=True)
            back["status"][~hit] = _abi.ATR_ST_NONE
            rp.rounds = (front, back)
            STATS["batched_gpu_calls"] += 2
            return
        aset = self._set_of(c, list(c.adapters), device)
        front_flags = np.array([-1 if a._front_flag is None else int(a._front_flag) for a in c.adapters])
        lo = np.zeros(n, dtype=np.int64)
        hi = lens.copy()
        active = hi > lo                          # `if len(read) == 0: return read` (modifiers.py:136-137)
        rounds = []
        for _ in range(c.times):                  # modifiers.py:143-149
            win = np.stack([lo, hi], axis=1).astype(np.uint16)
            res = aset.locate_host(ascii, offsets, win=win, fold_case=True)
            STATS["batched_gpu_calls"] += 1
            res["status"][~active] = _abi.ATR_ST_NONE
            hit = res["status"] == _abi.ATR_ST_MATCH
            rounds.append(res.tolist())
            if not hit.any():
                break
            ff = front_flags[np.clip(res["adapter"], 0, None)]
            is_front = np.where(ff < 0, res["rstart"] == 0, ff == 1)
            lo = np.where(hit & is_front, lo + res["rstop"], lo)      # Adapter._trimmed_front keeps read[rstop:]
            hi = np.where(hit & ~is_front, lo + res["rstart"], hi)     # Adapter._trimmed_back keeps read[:rstart]
            active = hit
        rp.rounds = rounds

    def _precompute_insert(self, cutter, state, device):
        from .align import InsertAligner
        rp = self.replays[0]
        ref_al = cutter.__dict__.get("_atr_ref_aligner")
        if ref_al is None:
            ref_al = cutter.__dict__["_atr_ref_aligner"] = cutter.aligner
        key = engine.context_key(device)
        gpu_al = cutter.__dict__.setdefault("_atr_aligners", {}).get(key)
        if gpu_al is None:
            gpu_al = cutter.__dict__["_atr_aligners"][key] = InsertAligner(
                ref_al.adapter1, ref_al.adapter2, match_probability=ref_al.match_probability,
                insert_max_rmp=ref_al.insert_max_rmp, adapter_max_rmp=ref_al.adapter_max_rmp,
                min_insert_overlap=ref_al.min_insert_overlap, max_insert_mismatch_frac=ref_al.max_insert_mismatch_frac,
                min_adapter_overlap=ref_al.min_adapter_overlap, max_adapter_mismatch_frac=ref_al.max_adapter_mismatch_frac,
                adapter_check_cutoff=ref_al.adapter_check_cutoff, base_probs=ref_al.base_probs,
                adapter_wildcards=ref_al.adapter_wildcards, read_wildcards=ref_al.read_wildcards, device=device)
        a1, o1 = engine.encode_reads([s[0].sequence for s in state])
        a2, o2 = engine.encode_reads([s[1].sequence for s in state])
        l1, l2 = np.diff(o1), np.diff(o2)
        ins = gpu_al.match_insert_batch((a1, o1), (a2, o2))
        STATS["batched_gpu_calls"] += 1
        skipped = (l1 < cutter.min_insert_len) | (l2 < cutter.min_insert_len)          # modifiers.py:392-394
        need = (ins["insert"]["status"] == _abi.ATR_ST_NONE) & ~skipped                # :401-406
        fb = []
        for adapter, (a, o, l) in zip((cutter.adapter1, cutter.adapter2), ((a1, o1, l1), (a2, o2, l2))):
            win = np.zeros((len(l), 2), dtype=np.uint16)
            win[need, 1] = l[need]
            rec = self._set_of(cutter, [adapter], device).locate_host(a, o, win=win, fold_case=True)
            STATS["batched_gpu_calls"] += 1
            rec["status"][~need] = _abi.ATR_ST_NONE
            fb.append(rec)
        rp.ins, rp.fb = ins, fb

    # -- swap the three call sites of the adapter stage for the replay, and back ----------------------------
    def __enter__(self):
        stage = self.mods.modifiers[self.index]
        if self.kind == "insert":
            rp = self.replays[0]
            stage.__dict__.setdefault("_atr_ref_aligner", stage.aligner)
            stage.aligner = rp
            stage.adapter1.match_to = rp.match_to(0)
            stage.adapter2.match_to = rp.match_to(1)
        else:
            for rp in self.replays:
                if rp is not None:
                    rp.cutter._best_match = rp
        return self

    def __exit__(self, *exc):
        stage = self.mods.modifiers[self.index]
        if self.kind == "insert":
            stage.aligner = stage.__dict__["_atr_ref_aligner"]
            del stage.adapter1.__dict__["match_to"]
            del stage.adapter2.__dict__["match_to"]
        else:
            for rp in self.replays:
                if rp is not None:
                    del rp.cutter.__dict__["_best_match"]
        return False

    def seek(self, idx):
        for rp in self.replays:
            if rp is not None:
                rp.seek(idx)


def _plan_for(mods):
    plan = mods.__dict__.get("_atr_plan")
    if plan is None or plan.n_mods != len(mods.modifiers):
        plan = _Plan(mods)
        plan.n_mods = len(mods.modifiers)
        mods.__dict__["_atr_plan"] = plan
    return plan if plan.kind is not None else None


def _batched_handle_records(self, context, records):
    """TrimPipeline.handle_records (commands/trim/__init__.py:82-84), staged around one GPU adapter stage per batch.
    Reproduces, per record and in order, Pipeline.handle_records (base.py:65-78), Single/PairedEndPipelineMixin.
    handle_record (:113-127), StatsRecordHandlerWrapper.handle_record (trim/__init__.py:163-172) and
    RecordHandler.handle_record (:121-126)."""
    atropos, _, _, base, trim, _ = _ref()
    handler = self.record_handler
    wrapper = None
    if isinstance(handler, trim.StatsRecordHandlerWrapper):
        wrapper, handler = handler, handler.record_handler
    plan = _plan_for(handler.modifiers) if isinstance(handler, trim.RecordHandler) else None
    if plan is None:
        STATS["percall_batches"] += 1
        return _orig_handle_records(self, context, records)
    paired = isinstance(self, base.PairedEndPipelineMixin)
    bps = context['bp']
    source = context['source']

    def failed(idx, err):
        raise atropos.AtroposError(
            "An error occurred at record {} of batch {}".format(idx, context['index'])) from err

    state = []
    for idx, record in enumerate(records):
        try:
            if paired:
                read1, read2 = record
                bps[0] += len(read1.sequence)
                bps[1] += len(read2.sequence)
            else:
                read1, read2 = record, None
                bps[0] += len(record)
            if wrapper is not None and wrapper.pre is not None:
                wrapper.collect(wrapper.pre, source, read1, read2, **wrapper.pre_kwargs)
            state.append(plan.before(read1, read2))
        except Exception as err:
            failed(idx, err)
    try:
        plan.precompute(state, device=_device())
    except Exception as err:
        failed(0, err)
    mids = None
    if plan.merge is not None:
        # two GPU stages: the adapter stage's replay runs over the whole batch first, then ONE merge-alignment call
        mids = []
        with plan:
            for idx, (read1, read2) in enumerate(state):
                try:
                    plan.seek(idx)
                    mids.append(plan.stage_to_merge(read1, read2))
                except _Desync:
                    raise
                except Exception as err:
                    failed(idx, err)
        try:
            plan.merge.precompute(mids, device=_device())
        except Exception as err:
            failed(0, err)
        STATS["batched_merge_batches"] = STATS.get("batched_merge_batches", 0) + 1
    with plan:
        for idx, (read1, read2) in enumerate(state if mids is None else mids):
            try:
                if mids is None:
                    plan.seek(idx)
                    reads = plan.from_stage(read1, read2)
                else:
                    plan.merge.seek(idx)
                    reads = plan.from_merge(read1, read2)
                dest = handler.filters.filter(*reads)
                handler.formatters.format(context["results"], dest, *reads)
                if wrapper is not None and wrapper.post is not None:
                    if dest not in wrapper.post:
                        wrapper.post[dest] = {}
                    wrapper.collect(wrapper.post[dest], source, *reads, **wrapper.post_kwargs)
            except _Desync:
                raise
            except Exception as err:
                failed(idx, err)
    self.result_handler.write_result(context["index"], context["results"])


def _device():
    import os
    return int(os.environ.get("ATROPOS_B200_DEVICE", "0"))


def install_batched_pipeline():
    global _orig_handle_records
    if _installed["batched"]:
        return
    if not _installed["finder"]:
        install(batched=False)
    import atropos.commands.trim as trim
    if not shim_active():
        raise RuntimeError("atropos.align._align is not the atropos_b200 module")
    _orig_handle_records = trim.TrimPipeline.handle_records
    trim.TrimPipeline.handle_records = _batched_handle_records
    _installed["batched"] = True


def main(argv=None):
    """``python -m atropos_b200.integration [--aligner-backend gpu|gpu-per-call|cython] <atropos command line>``: the
    reference's own launcher (bin/atropos -> atropos.commands.execute_cli) on top of this engine."""
    argv = list(sys.argv[1:] if argv is None else argv)
    # backend switch ("--aligner-backend {cython,gpu}-style switch / env var in the shim"):
    #   --aligner-backend gpu | gpu-per-call | cython   or   ATROPOS_ALIGNER_BACKEND=...   (flag wins; default gpu)
    # `cython` leaves the reference exactly as it is (its own compiled _align module): the A/B partner of the other two.
    import os
    backend = os.environ.get("ATROPOS_ALIGNER_BACKEND", "gpu")
    while argv and argv[0] in ("--per-call", "--aligner-backend"):
        if argv[0] == "--per-call":
            backend, argv = "gpu-per-call", argv[1:]
        else:
            if len(argv) < 2:
                raise SystemExit("--aligner-backend needs a value: gpu, gpu-per-call or cython")
            backend, argv = argv[1], argv[2:]
    if backend not in ("gpu", "gpu-per-call", "cython"):
        raise SystemExit("unknown aligner backend %r (gpu, gpu-per-call, cython)" % (backend,))
    if backend != "cython":
        install(batched=(backend == "gpu"))
    from atropos.commands import execute_cli
    return execute_cli(argv)


if __name__ == "__main__":
    sys.exit(main())
