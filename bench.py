#!/usr/bin/env python
"""bench.py -- headline benchmark of the adapter-alignment hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N ...            # the reference's own CPU implementation

Workload (config.workload = "cfg2"): SE 150 bp, 10 M synthetic reads PER GPU, one 3' TruSeq adapter
(34 nt), max error rate 0.1, min overlap 3 -- `Adapter.match_to` for every read. A "step" is one pass
of the hot path over the whole batch.

  value  M reads/s with the packed reads already resident in HBM (kernel(s) only, CUDA events on the
         launching stream, barrier + synchronize on both sides, max over ranks);
  e2e    the same metric through the host entry point `atr_locate_batch_host` (what the Python
         `Adapter.match_to_batch` calls): ASCII reads in pinned host memory -> H2D -> pack -> align ->
         D2H of the 16-byte records, all inside the timed region;
  roofline  algorithmic bytes (95 B/read: 75 packed + 4 offset + 16 result) / measured kernel time,
         against the measured HBM copy bandwidth in MEASURED_PEAKS.json;
  cpu_baseline  the reference's compiled Cython aligner (oracle/_ref) on a bounded sample on this
         box's host cores (rank 0, N=1 only).

Multi-GPU (torchrun, one rank per GPU): reads are independent -> every rank aligns its own shard
(seed + rank), no collective on the data path; "scaling": "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

READ_LEN = 150
READS_PER_GPU = 10_000_000
ADAPTER = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"
ERROR_RATE = 0.1
MIN_OVERLAP = 3
ALGO_BYTES_PER_READ = (READ_LEN + 1) // 2 + 4 + 16          # 95 B @ L=150
CONFIG = {"workload": "cfg2: SE 150 bp, 10M synthetic reads per GPU, 3' TruSeq adapter (34 nt), err 0.1, overlap 3",
          "reads_per_gpu": READS_PER_GPU, "read_len": READ_LEN, "adapter_len": len(ADAPTER),
          "l2_policy": "inputs larger than L2 (>= 950 MB per step vs 126 MB L2)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.marks = index, [], None, []

    def mark(self):
        """remember how many samples existed at this point (start / end of the kernel-only timed region)"""
        self.marks.append(len(self.rows))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = self.rows
        timed = rows[self.marks[0]:self.marks[1] + 1] if len(self.marks) >= 2 else []
        under_load = timed if len(timed) >= 2 else rows      # short timed regions: fall back to warm-up + timed + e2e
        sm = sorted(int(float(r[1])) for r in under_load if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(timed)}


# ------------------------------------------------------------------------------------------------
# reference / cpu baseline: the reference's own compiled aligner, one process per host core
# ------------------------------------------------------------------------------------------------
def _ref_worker(args):
    blob, n, L = args
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    if ref_loader.native_available():
        aligner = ref_loader.load_native().Aligner(ADAPTER, ERROR_RATE, 14, False, False, MIN_OVERLAP, 1)
        loc = aligner.locate
    else:                                   # oracle port (C restatement) -- only if oracle/_ref did not travel
        from oracle import oracle as orc
        loc = lambda q: orc.locate(ADAPTER, q, ERROR_RATE, 14, False, False, MIN_OVERLAP, 1)  # noqa: E731
    text = blob.decode("ascii")
    t0 = time.perf_counter()
    hits = 0
    for i in range(n):
        if loc(text[i * L:(i + 1) * L]) is not None:
            hits += 1
    return time.perf_counter() - t0, hits


def cpu_reference_rate(reads_cpu, sample, cores):
    """Time the reference's Aligner.locate over `sample` reads split over `cores` processes.
    Returns (M reads/s, seconds, kind)."""
    import multiprocessing as mp
    from oracle import ref_loader
    kind = "reference" if ref_loader.native_available() else "port"
    sample = min(sample, reads_cpu.shape[0])
    per = sample // cores
    chunks = [(bytes(reads_cpu[c * per:(c + 1) * per].reshape(-1)), per, READ_LEN) for c in range(cores)]
    if cores == 1:
        dt, _ = _ref_worker(chunks[0])
        return per / dt / 1e6, dt, kind
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_worker, [(b"", 0, READ_LEN)] * cores)          # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(_ref_worker, chunks)
        wall = time.perf_counter() - t0
    return per * cores / wall / 1e6, wall, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from atropos_b200 import synth
    cores = os.cpu_count() or 1
    sample_per_step = 60_000 * cores                   # ~0.5-1 s per core per step at ~0.1 M reads/s/core
    reads = synth.synth_se(sample_per_step, READ_LEN, ADAPTER, seed=synth.seed_for(2), device="cpu").numpy()
    rates = []
    for s in range(args.warmup + args.steps):
        r, dt, kind = cpu_reference_rate(reads, sample_per_step, cores)
        if s >= args.warmup:
            rates.append((r, dt))
    tot_t = sum(dt for _, dt in rates)
    value = sample_per_step * len(rates) / tot_t / 1e6
    line = {"impl": "reference", "metric": "M reads/sec trimmed (150 bp SE, TruSeq 3' adapter, err 0.1)",
            "value": value, "unit": "M reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_t / len(rates) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": dict(CONFIG, sample_per_step=sample_per_step),
            "cpu_baseline": {"value": value, "unit": "M reads/s", "cores": cores, "kind": kind,
                             "sample": "%d reads per step, Aligner.locate of the reference's compiled Cython module, "
                                       "one process per host core" % sample_per_step},
            "e2e": {"value": value, "unit": "M reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local):
    """Run this rank (and therefore place its pinned host buffers: pages are pinned on the node of the allocating CPU)
    on the NUMA node its GPU hangs off. With several ranks on one box the copies otherwise cross the socket
    interconnect: 4 ranks reached only 2.0x the single-GPU end-to-end rate. Returns a description for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as exc:                      # no sysfs / no such attribute: leave the placement to the OS
        return "not bound: %r" % (exc,)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from atropos_b200 import _abi, engine, synth
    from atropos_b200.adapters import Adapter, BACK

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_binding = bind_to_gpu_numa_node(torch, local) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()                      # create the NCCL communicator before the big allocations
        torch.cuda.synchronize()

    n, L = args.reads or READS_PER_GPU, READ_LEN
    # ---- synthetic shard, generated in HBM; a pinned host copy feeds the e2e leg -----------------
    reads_dev = synth.synth_se(n, L, ADAPTER, seed=synth.seed_for(2, rank), device=dev)          # uint8 [n, L]
    offsets_dev = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    reads_host = torch.empty((n, L), dtype=torch.uint8, pin_memory=True)
    reads_host.copy_(reads_dev)
    offsets_host = torch.empty(n + 1, dtype=torch.int64, pin_memory=True)
    offsets_host.copy_(offsets_dev)
    out_host = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    ctx = engine.default_context(local)
    adapter = Adapter(ADAPTER, BACK, max_error_rate=ERROR_RATE, min_overlap=MIN_OVERLAP, device=local)
    aset = adapter._adapterset()
    assert aset.ctx is ctx
    L_ = ctx._L

    # ---- pack once: the HBM-resident layout the kernel metric is quoted on ------------------------
    words = n * ((L + 7) // 8)
    codes = torch.empty(words + 8, dtype=torch.int32, device=dev)
    woff = torch.empty(n + 1, dtype=torch.int32, device=dev)
    lens = torch.empty(n, dtype=torch.int16, device=dev)
    out_dev = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    engine._lib.check(L_.atr_pack_device(ctx.handle, reads_dev.data_ptr(), offsets_dev.data_ptr(), n, 1,
                                         codes.data_ptr(), woff.data_ptr(), lens.data_ptr()), ctx.handle)
    ctx.sync()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step_device():
        aset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out_dev.data_ptr())

    def step_host():
        aset.locate_host(reads_host.numpy().reshape(-1), offsets_host.numpy(), fold_case=True,
                         out=out_host.numpy().view(_abi.MATCH_DTYPE).reshape(-1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel-only leg (value + roofline) ---------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)                      # let nvidia-smi start streaming before the timed region
    for _ in range(args.warmup):
        step_device()
    ctx.sync()
    barrier()
    if rank == 0:
        sampler.mark()
    ctx.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    ev1.synchronize()
    barrier()
    launches = ctx.launch_count()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    if rank == 0:
        sampler.mark()
    ms_per_step = dev_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---- per-kernel times of the same step (CUDA events between the launches, separate pass) -------------
    ctx.set_profiling(True)
    phase = []
    for _ in range(args.steps):
        step_device()
        ph = ctx.last_phase_ms()
        if len(ph) == 4:
            phase.append(ph)
    ctx.set_profiling(False)
    ctx.sync()
    kernels_ms = None
    if phase:
        names = ["k_filter_sa", "k_refine", "k_band", "k_wide"]
        kernels_ms = {nm: sum(p[i] for p in phase) / len(phase) for i, nm in enumerate(names)}

    # ---- e2e leg: host buffers through the public host entry point ------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(min(args.warmup, 2) or 1):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e_value = world * n / e2e_s / 1e6

    # ---- FASTQ leg ("next" rows f-1/f-2/f-3): FASTQ text in pinned host memory -> atr_trim_fastq_host -> trimmed
    # FASTQ text + the report's statistics in host memory. Same reads; reader, trimming and formatting on the GPU.
    fq = None
    if not args.no_fastq:
        from atropos_b200 import fastq as fastq_mod
        n_fq = n if world == 1 else min(n, 4_000_000)           # several ranks share host memory and PCIe switches
        text_np = synth.fastq_text(reads_host.numpy()[:n_fq])
        text_host = torch.empty(text_np.size, dtype=torch.uint8, pin_memory=True)
        text_host.numpy()[:] = text_np
        del text_np
        fq_out = torch.empty(text_host.numel(), dtype=torch.uint8, pin_memory=True)
        trimmer = fastq_mod.FastqTrimmer([adapter], times=1, max_len=L, device=local)
        fq_steps = max(1, min(args.steps, 3))
        res = trimmer.trim(text_host.numpy(), out=fq_out.numpy())          # warm-up (allocations)
        barrier()
        ctx.launch_count(reset=True)
        t0 = time.perf_counter()
        for _ in range(fq_steps):
            res = trimmer.trim(text_host.numpy(), out=fq_out.numpy())
        barrier()
        fq_s = max_over_ranks(time.perf_counter() - t0) / fq_steps
        fq_launches = ctx.launch_count() // fq_steps
        out_view, fq_stats, _ = res
        assert fq_stats.records == n_fq
        fq = {"value": world * n_fq / fq_s / 1e6, "reads_per_gpu": n_fq, "unit": "M reads/s", "ms_per_step": fq_s * 1e3, "steps": fq_steps,
              "h2d_bytes_per_step": int(text_host.numel()), "d2h_bytes_per_step": int(out_view.size),
              "gpu_launches_per_step": int(fq_launches), "reads_with_adapters": int(fq_stats.with_adapters),
              "api": "atr_trim_fastq_host (fastq.FastqTrimmer.trim): FASTQ text -> trimmed FASTQ text + report statistics"}
    clocks = sampler.stop() if rank == 0 else None

    # parity spot check of the timed outputs (device leg vs host leg must agree bit for bit)
    a = out_dev.cpu().numpy().view(_abi.MATCH_DTYPE).reshape(-1)
    b = out_host.numpy().view(_abi.MATCH_DTYPE).reshape(-1)
    assert np.array_equal(a, b), "device-resident and host entry points disagree"
    hit_frac = float((a["status"] == _abi.ATR_ST_MATCH).mean())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = load_peaks()
    step_achieved = ALGO_BYTES_PER_READ * n / (ms_per_step * 1e-3) / 1e9     # per GPU, whole step (all kernels of the path)
    # dominant kernel: algorithmic bytes one launch processes / that kernel's average launch duration
    dom_name, dom_ms = ("step", ms_per_step)
    if kernels_ms:
        dom_name = max(kernels_ms, key=kernels_ms.get)
        dom_ms = kernels_ms[dom_name]
    achieved = ALGO_BYTES_PER_READ * n / (dom_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(dom_name, {}).get("dram_bytes_per_read", None)
            traffic = None if traffic is None else traffic * n
        except Exception:
            traffic = None
    line = {
        "metric": "M reads/sec trimmed (150 bp SE, TruSeq 3' adapter, err 0.1)",
        "value": value, "unit": "M reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 packed integer DP (4-bit bases)", "data": "synthetic",
        "config": dict(CONFIG, reads_per_gpu=n, adapter_hit_fraction=round(hit_frac, 4), host_binding=host_binding),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "M reads/s", "h2d_bytes_per_step": int(n * L),      # fixed-length batch: the offsets are rebuilt on the device
                "d2h_bytes_per_step": int(16 * n), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "api": "atr_locate_batch_host (Adapter.match_to_batch)"},
        "e2e_fastq": fq,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": dom_name, "kernel_ms": dom_ms,
                     "kernel_share_of_step": (dom_ms / sum(kernels_ms.values())) if kernels_ms else 1.0,
                     "kernels_ms": kernels_ms, "algorithmic_bytes_per_read": ALGO_BYTES_PER_READ,
                     "step_achieved": step_achieved, "step_frac": step_achieved / peak,
                     "note": "integer-ALU bound path: see DESIGN.md; equivalent full-matrix cell rate below",
                     "gcups_equivalent": n * READ_LEN * len(ADAPTER) / (ms_per_step * 1e-3) / 1e9},
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(n, 400_000 * cores)
        rate, secs, kind = cpu_reference_rate(reads_host.numpy(), sample, cores)
        r1, s1, _ = cpu_reference_rate(reads_host.numpy(), 300_000, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "M reads/s", "cores": cores, "kind": kind,
                                "sample": "first %d reads of the same batch, reference Aligner.locate (compiled Cython), "
                                          "one process per core, %.1f s; single core: %.3f M reads/s" % (sample, secs, r1)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the 10 M of BASELINE config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fastq", action="store_true", help="skip the FASTQ-text leg (e2e_fastq)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
