#!/usr/bin/env python
"""bench.py -- headline benchmark of the adapter-alignment hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N ...            # the reference's own CPU implementation

Headline workload (config.workload = "cfg2"): SE 150 bp, 10 M synthetic reads PER GPU, one 3' TruSeq adapter
(34 nt), max error rate 0.1, min overlap 3 -- `Adapter.match_to` for every read. A "step" is one pass of the hot path
over the whole batch.

  value     M reads/s with the packed reads already resident in HBM (kernel(s) only, CUDA events on the launching
            stream, barrier + synchronize on both sides, max over ranks);
  e2e       the same metric through the host entry point `atr_locate_batch_host` (what the Python
            `Adapter.match_to_batch` calls): ASCII reads in pinned host memory -> H2D -> pack -> align -> D2H of the
            16-byte records, all inside the timed region; `e2e.link_roof` = the same bytes over the same link with no
            kernel at all (what the box's PCIe / host memory path allows at this N);
  roofline  algorithmic bytes (95 B/read: 75 packed + 4 offset + 16 result) / measured time of the dominant kernel,
            against the measured HBM copy bandwidth in MEASURED_PEAKS.json;
  cpu_baseline  the reference's compiled Cython aligner (oracle/_ref) on a bounded sample on this box's host cores
            (rank 0, N=1 only), plus the reference's own `atropos trim -T N` command line on the same bytes;
  strong_scaling  cfg 2 with the 10 M reads split over the N ranks (the weak-scaling `value` keeps 10 M per GPU);
  configs   the other BASELINE configurations at their named per-GPU scale: cfg3 (PE 2x150 insert aligner, 10 M
            pairs per GPU), cfg4 (8-adapter panel, 25 M reads per GPU = 100 M over 4 GPUs), cfg5 (PE 2x300, error
            rate 0.15, streamed in 4 M-pair chunks: --cfg5-chunks 31 per GPU at 8 GPUs is the 1 B-pair job), each with
            value / roofline / e2e / hit fraction and a >= 100 k-unit check of the timed output against the oracle.

Same bytes for both arms: the first 2 Mi reads of shard 0 come from the CPU generator (torch CPU != CUDA random
streams) and are what `--impl reference` and `cpu_baseline` time; the rest of the batch is generated in HBM.

Multi-GPU (torchrun, one rank per GPU): reads are independent -> every rank aligns its own shard (seed + rank), no
collective on the data path; "scaling": "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
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
CPU_PREFIX = 2 << 20                                         # reads of shard 0 that come from the CPU generator
ALGO_BYTES_PER_READ = (READ_LEN + 1) // 2 + 4 + 16          # 95 B @ L=150
CONFIG = {"workload": "cfg2: SE 150 bp, 10M synthetic reads per GPU, 3' TruSeq adapter (34 nt), err 0.1, overlap 3",
          "reads_per_gpu": READS_PER_GPU, "read_len": READ_LEN, "adapter_len": len(ADAPTER),
          "l2_policy": "inputs larger than L2 (>= 950 MB per step vs 126 MB L2)"}
# cfg 4: the panel of SURVEY 8d (three 3' adapters, two anchored 5', one unanchored 5'; the two linked adapters are
# represented by their components: the reference itself cannot rank a LinkedMatch inside a panel, modifiers.py:120)
PANEL = [("AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC", "BACK"),
         ("AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT", "BACK"),
         ("TGGAATTCTCGGGTGCCAAGG", "BACK"),
         ("GTTCAGAGTTCTACAGTCCGACGATC", "PREFIX"),
         ("ACACTCTTTCCCTACACGACGCTCTTCCGATCT", "PREFIX"),
         ("AATGATACGGCGACCACCGA", "FRONT"),
         ("TGGAATTCTCGGGTGCCAAGG", "BACK"),
         ("AGATCGGAAGAGC", "BACK")]


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
        self.nvml_rows, self.nvml_on, self.nvml_thread = [], False, None

    def mark(self):
        """remember how many samples existed at this point (start / end of the kernel-only timed region). Between the
        two marks an in-process NVML poller (same counters as nvidia-smi's: SM clock, clocks-event reasons) samples every
        ~1 ms: nvidia-smi's 50 ms period puts 0-1 samples into a 25 ms timed region."""
        self.marks.append(len(self.rows))
        if len(self.marks) == 1:
            self._nvml_start()
        elif len(self.marks) == 2:
            self.nvml_on = False
            if self.nvml_thread is not None:
                self.nvml_thread.join(timeout=2)

    def _nvml_start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return
        self.nvml_on = True

        def poll():
            while self.nvml_on and len(self.nvml_rows) < 100000:
                try:
                    self.nvml_rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                           pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
                except Exception:
                    break
                time.sleep(0.001)
        self.nvml_thread = threading.Thread(target=poll, daemon=True)
        self.nvml_thread.start()

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
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(timed)}
        if self.nvml_rows:
            # NVML reason bits (nvml.h nvmlClocksEventReasons*): 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal, 0x40 hw_thermal
            bits = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
            clk = sorted(c for c, _ in self.nvml_rows)
            seen = set()
            for _, r in self.nvml_rows:
                for b, nm in bits.items():
                    if r & b:
                        seen.add(nm)
            out["timed_region_nvml"] = {"samples": len(clk), "sm_mhz_median": clk[len(clk) // 2], "sm_mhz_min": clk[0],
                                        "reasons": sorted(seen)}
            if len(timed) < 2:
                out["sm_mhz"] = clk[len(clk) // 2]
            out["reasons"] = sorted(set(out["reasons"]) | seen)
        return out


# ------------------------------------------------------------------------------------------------
# synthetic input
# ------------------------------------------------------------------------------------------------
def cpu_prefix_reads(n):
    """The first reads of shard 0, from the CPU generator: the bytes both arms (and cpu_baseline) look at."""
    from atropos_b200 import synth
    return synth.synth_se(min(n, CPU_PREFIX), READ_LEN, ADAPTER, seed=synth.seed_for(2), device="cpu").numpy()


# ------------------------------------------------------------------------------------------------
# reference / cpu baseline: the reference's own compiled aligner, one process per host core; the reference's own CLI
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


def reference_cli_rates(reads_cpu, cores, serial_reads=100_000, budget_s=120.0):
    """`atropos trim` of the UNMODIFIED reference package (baseline/_ref, staged by oracle/build_ref.py) on the same
    reads as FASTQ: -T N with the writer process, -T N --no-writer-process, and serial (no -T: the CLI rejects -T 1,
    commands/cli.py:538-539). The parallel mode has a fixed >= 5 s poll latency (multicore.py:13, 344-350), hence
    >= 2 M reads. Returns a dict of M reads/s (None where a run failed or the budget ran out)."""
    from atropos_b200 import synth
    stage = os.path.join(ROOT, "baseline", "_ref")
    launcher = os.path.join(stage, "bin", "atropos")
    if not os.path.exists(launcher):
        return {"unavailable": "baseline/_ref not staged (python oracle/build_ref.py in the build container)"}
    out = {"reads": int(reads_cpu.shape[0]), "threads": cores,
           "command": "atropos trim -T N -a %s -e %g --no-default-adapters --no-cache-adapters --quiet "
                      "--report-file /dev/null -o out.fq -se in.fq" % (ADAPTER, ERROR_RATE)}
    tmp = tempfile.mkdtemp(prefix="atr_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    env = dict(os.environ, PYTHONPATH=stage + os.pathsep + os.environ.get("PYTHONPATH", ""))
    t_begin = time.perf_counter()
    try:
        fq = os.path.join(tmp, "in.fq")
        synth.fastq_text(reads_cpu).tofile(fq)
        fq_small = os.path.join(tmp, "small.fq")
        synth.fastq_text(reads_cpu[:serial_reads]).tofile(fq_small)
        base = [sys.executable, launcher, "trim", "-a", ADAPTER, "-e", str(ERROR_RATE), "--no-default-adapters",
                "--no-cache-adapters", "--quiet", "--report-file", "/dev/null", "-o", os.path.join(tmp, "out.fq")]
        runs = [("trim_threads", ["-T", str(cores)], fq, reads_cpu.shape[0]),
                ("trim_threads_no_writer", ["-T", str(cores), "--no-writer-process"], fq, reads_cpu.shape[0]),
                ("trim_serial", [], fq_small, min(serial_reads, reads_cpu.shape[0]))]
        for key, extra, path, nreads in runs:
            left = budget_s - (time.perf_counter() - t_begin)
            if left < 10 or (cores < 2 and extra):
                out[key] = None
                continue
            t0 = time.perf_counter()
            try:
                p = subprocess.run(base + extra + ["-se", path], env=env, cwd=tmp, stdout=subprocess.DEVNULL,
                                   stderr=subprocess.PIPE, timeout=left)
                dt = time.perf_counter() - t0
                out[key] = (nreads / dt / 1e6) if p.returncode == 0 else None
                if p.returncode != 0:
                    out[key + "_error"] = p.stderr.decode("latin-1")[-300:]
            except subprocess.TimeoutExpired:
                out[key] = None
                out[key + "_error"] = "time budget of the baseline exhausted"
    finally:
        for f in os.listdir(tmp):
            try:
                os.remove(os.path.join(tmp, f))
            except OSError:
                pass
        try:
            os.rmdir(tmp)
        except OSError:
            pass
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample_per_step = min(60_000 * cores, CPU_PREFIX)  # ~0.5-1 s per core per step at ~0.1 M reads/s/core
    reads = cpu_prefix_reads(CPU_PREFIX)               # the very bytes the GPU arm's shard 0 starts with
    rates = []
    for s in range(args.warmup + args.steps):
        r, dt, kind = cpu_reference_rate(reads, sample_per_step, cores)
        if s >= args.warmup:
            rates.append((r, dt))
    tot_t = sum(dt for _, dt in rates)
    value = sample_per_step * len(rates) / tot_t / 1e6
    cli = {} if args.no_cli_baseline else reference_cli_rates(reads, cores)
    line = {"impl": "reference", "metric": "M reads/sec trimmed (150 bp SE, TruSeq 3' adapter, err 0.1)",
            "value": value, "unit": "M reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_t / len(rates) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": dict(CONFIG, sample_per_step=sample_per_step),
            "cpu_baseline": {"value": value, "unit": "M reads/s", "cores": cores, "kind": kind,
                             "sample": "%d reads per step = the first reads of the GPU arm's shard 0 (same bytes), "
                                       "Aligner.locate of the reference's compiled Cython module, one process per "
                                       "host core" % sample_per_step,
                             "reference_cli": cli},
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


class Rig(object):
    """What every measurement of our arm needs: the rank layout, the engine context, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from atropos_b200 import engine
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.host_binding = bind_to_gpu_numa_node(torch, self.local) if self.world > 1 else "single rank: not bound"
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            dist.barrier()                  # create the NCCL communicator before the big allocations
            torch.cuda.synchronize()
        self.ctx = engine.default_context(self.local)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        self.peak, self.peak_src = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def time_device(self, fn, steps, warmup):
        """ms per step of `fn` (launches on the ctx stream): CUDA events on that stream, barrier + synchronize on both
        sides, max over ranks. Also returns the kernels launched inside the timed region."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.ctx.sync()
        self.barrier()
        self.ctx.launch_count(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        for _ in range(steps):
            fn()
        ev1.record(self.stream)
        ev1.synchronize()
        self.barrier()
        return self.max_over_ranks(ev0.elapsed_time(ev1)) / steps, self.ctx.launch_count()

    def time_host(self, fn, steps, warmup):
        """seconds per step of a blocking host call (wall clock around barrier-bracketed calls, max over ranks)"""
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return self.max_over_ranks(time.perf_counter() - t0) / steps

    def pack(self, reads_dev, fold_case):
        """ASCII reads [n, L] in HBM -> the packed layout (codes, woff, len, offsets)"""
        torch, engine_lib = self.torch, self.ctx._L
        from atropos_b200 import engine
        n, L = reads_dev.shape
        offs = torch.arange(n + 1, dtype=torch.int64, device=self.dev) * L
        codes = torch.empty(n * ((L + 7) // 8) + 8, dtype=torch.int32, device=self.dev)
        woff = torch.empty(n + 1, dtype=torch.int32, device=self.dev)
        lens = torch.empty(n, dtype=torch.int16, device=self.dev)
        torch.cuda.synchronize()            # the ctx stream does not synchronise with torch's streams
        engine._lib.check(engine_lib.atr_pack_device(self.ctx.handle, reads_dev.data_ptr(), offs.data_ptr(), n,
                                                     int(fold_case), codes.data_ptr(), woff.data_ptr(), lens.data_ptr()),
                          self.ctx.handle)
        self.ctx.sync()
        return codes, woff, lens, offs

    def pinned_copy(self, t):
        h = self.torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        self.torch.cuda.synchronize()
        return h

    def link_roof(self, h2d_bytes, d2h_bytes, steps=3):
        """The same bytes over the same link with no kernel: pinned host -> HBM on one stream, HBM -> pinned host on
        another, concurrently (what a perfectly overlapped pipeline could reach on this box at this N)."""
        torch = self.torch
        chunk = 256 << 20
        src = torch.empty(min(h2d_bytes, chunk), dtype=torch.uint8, pin_memory=True)
        dst = torch.empty(src.numel(), dtype=torch.uint8, device=self.dev)
        osrc = torch.empty(min(max(d2h_bytes, 1), chunk), dtype=torch.uint8, device=self.dev)
        odst = torch.empty(osrc.numel(), dtype=torch.uint8, pin_memory=True)
        s1, s2 = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)

        def once():
            done = 0
            with torch.cuda.stream(s1):
                while done < h2d_bytes:
                    k = min(src.numel(), h2d_bytes - done)
                    dst[:k].copy_(src[:k], non_blocking=True)
                    done += k
            done = 0
            with torch.cuda.stream(s2):
                while done < d2h_bytes:
                    k = min(osrc.numel(), d2h_bytes - done)
                    odst[:k].copy_(osrc[:k], non_blocking=True)
                    done += k
            s1.synchronize()
            s2.synchronize()
        sec = self.time_host(once, steps, 1)
        return {"seconds_per_step": sec, "h2d_gbs": h2d_bytes / sec / 1e9, "d2h_gbs": d2h_bytes / sec / 1e9,
                "aggregate_h2d_gbs": self.world * h2d_bytes / sec / 1e9,
                "how": "pinned host <-> HBM copies of the same byte counts on two streams, no kernels, max over ranks"}


def roofline_block(rig, algo_bytes_per_unit, units, dom_name, dom_ms, kernels_ms, step_ms, extra=None):
    achieved = algo_bytes_per_unit * units / (dom_ms * 1e-3) / 1e9
    step_achieved = algo_bytes_per_unit * units / (step_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            ent = tj.get(dom_name.split("<")[0].split("+")[0], None)
            if ent is not None and ent.get("dram_bytes_per_read") is not None:
                traffic = ent["dram_bytes_per_read"] * units
                traffic_src = "profiles/traffic.json: ncu --set full capture '%s' of this kernel (dram__bytes_read.sum + " \
                              "dram__bytes_write.sum per launch), scaled to this launch; not measured in this run" % ent.get("capture", "?")
        except Exception:
            traffic = None
    rb = {"bound": "hbm", "achieved": achieved, "peak": rig.peak, "unit": "GB/s", "frac": achieved / rig.peak,
          "traffic": traffic, "traffic_source": traffic_src, "peak_source": rig.peak_src, "kernel": dom_name,
          "kernel_ms": dom_ms, "kernel_share_of_step": (dom_ms / sum(kernels_ms.values())) if kernels_ms else 1.0,
          "kernels_ms": kernels_ms, "algorithmic_bytes_per_unit": algo_bytes_per_unit,
          "step_achieved": step_achieved, "step_frac": step_achieved / rig.peak}
    if extra:
        rb.update(extra)
    return rb


# ------------------------------------------------------------------------------------------------
# cfg 2 (the headline)
# ------------------------------------------------------------------------------------------------
def run_cfg2(rig, args):
    import numpy as np
    torch = rig.torch
    from atropos_b200 import _abi, synth
    from atropos_b200.adapters import Adapter, BACK
    n, L = args.reads or READS_PER_GPU, READ_LEN
    dev, ctx, world, rank = rig.dev, rig.ctx, rig.world, rig.rank
    # ---- synthetic shard, generated in HBM; shard 0 starts with the CPU generator's reads (the reference arm's bytes) ----
    reads_dev = synth.synth_se(n, L, ADAPTER, seed=synth.seed_for(2, rank), device=dev)          # uint8 [n, L]
    prefix = None
    if rank == 0:
        prefix = cpu_prefix_reads(n)
        reads_dev[:prefix.shape[0]].copy_(torch.from_numpy(prefix))
    offsets_dev = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    reads_host = rig.pinned_copy(reads_dev)
    offsets_host = rig.pinned_copy(offsets_dev)
    out_host = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)

    adapter = Adapter(ADAPTER, BACK, max_error_rate=ERROR_RATE, min_overlap=MIN_OVERLAP, device=rig.local)
    aset = adapter._adapterset()
    assert aset.ctx is ctx
    codes, woff, lens, _ = rig.pack(reads_dev, True)
    out_dev = torch.empty((n, 16), dtype=torch.uint8, device=dev)

    def step_device():
        aset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out_dev.data_ptr())

    def step_host():
        aset.locate_host(reads_host.numpy().reshape(-1), offsets_host.numpy(), fold_case=True,
                         out=out_host.numpy().view(_abi.MATCH_DTYPE).reshape(-1))

    # ---- kernel-only leg (value + roofline) ---------------------------------------------------------
    sampler = ClockSampler(rig.local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)                      # let nvidia-smi start streaming before the timed region
    for _ in range(args.warmup):
        step_device()
    ctx.sync()
    rig.barrier()
    if rank == 0:
        sampler.mark()
    ms_per_step, launches = rig.time_device(step_device, args.steps, 0)
    if rank == 0:
        sampler.mark()
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---- per-kernel times of the same step (CUDA events between the launches, separate serialised pass) -------------
    ctx.set_profiling(True)
    phase, names = [], None
    for _ in range(min(args.steps, 20)):
        step_device()
        ph = ctx.last_phase_ms()
        if len(ph) == 4:
            phase.append(ph)
            names = ctx.last_phase_names()
    ctx.set_profiling(False)
    ctx.sync()
    kernels_ms = None
    if phase:
        kernels_ms = {nm: sum(p[i] for p in phase) / len(phase) for i, nm in enumerate(names) if nm}

    # ---- strong scaling: the 10 M reads of the configuration split over the ranks ------------------------------------
    strong = None
    if world > 1:
        ns = (args.reads or READS_PER_GPU) // world
        ms_s, _ = rig.time_device(lambda: aset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), ns, out_dev.data_ptr()),
                                  args.steps, 2)
        strong = {"value": world * ns / (ms_s * 1e-3) / 1e6, "unit": "M reads/s", "reads_total": world * ns,
                  "reads_per_gpu": ns, "ms_per_step": ms_s,
                  "note": "cfg 2's 10 M reads split evenly over the ranks (kernel time, max over ranks)"}
        step_device()                        # the full batch again: out_dev is compared with the host leg below
        ctx.sync()

    # ---- e2e leg: host buffers through the public host entry point ------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    e2e_s = rig.time_host(step_host, e2e_steps, min(args.warmup, 2) or 1)
    e2e_value = world * n / e2e_s / 1e6
    link = rig.link_roof(int(n * L), int(16 * n))
    link["M_reads_per_s"] = world * n / link["seconds_per_step"] / 1e6
    link["e2e_fraction_of_link_roof"] = e2e_value / link["M_reads_per_s"]

    # ---- e2e for callers that keep reads PACKED on the host (SURVEY 8d timing (ii)): pinned packed buffers -> H2D of the
    # codes only (75 B/read; fixed-length index rebuilt on the device) -> kernels -> D2H of the records; the host packer
    # (threaded AVX2, atr_pack_reads_host) is timed separately ----
    from atropos_b200 import engine as engine_mod
    nwords = n * ((L + 7) // 8)
    pk = [torch.empty((nwords + 8) * 4, dtype=torch.uint8, pin_memory=True), torch.empty((n + 1) * 4, dtype=torch.uint8, pin_memory=True),
          torch.empty(n * 2, dtype=torch.uint8, pin_memory=True)]
    pk_np = (pk[0].numpy().view(np.uint32), pk[1].numpy().view(np.uint32), pk[2].numpy().view(np.uint16))
    pack_s = rig.time_host(lambda: engine_mod.pack_reads_host(reads_host.numpy().reshape(-1), offsets_host.numpy(), fold_case=True,
                                                              out=pk_np), 2, 1)
    out_packed = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
    opv = out_packed.numpy().view(_abi.MATCH_DTYPE).reshape(-1)
    packed_s = rig.time_host(lambda: aset.locate_host_packed(pk_np[0], pk_np[1], pk_np[2], fold_case=True, out=opv), e2e_steps, 1)
    assert np.array_equal(opv, out_host.numpy().view(_abi.MATCH_DTYPE).reshape(-1)), "packed and ASCII host entry points disagree"
    plink = rig.link_roof(int(nwords * 4), int(16 * n))
    e2e_packed = {"value": world * n / packed_s / 1e6, "unit": "M reads/s", "h2d_bytes_per_step": int(nwords * 4),
                  "d2h_bytes_per_step": int(16 * n), "ms_per_step": packed_s * 1e3,
                  "api": "atr_locate_batch_host_packed: reads packed in pinned host memory (4-bit codes), records back",
                  "link_roof": dict(plink, M_reads_per_s=world * n / plink["seconds_per_step"] / 1e6),
                  "host_packer": {"M_reads_per_s": world * n / pack_s / 1e6, "ms": pack_s * 1e3, "threads": os.cpu_count(),
                                  "api": "atr_pack_reads_host (AVX2, one thread per core), ASCII -> packed, both in pinned host memory"},
                  "including_host_packing": {"value": world * n / (pack_s + packed_s) / 1e6,
                                             "note": "packer and packed call back to back, not overlapped"}}
    del pk, out_packed

    # ---- FASTQ leg ("next" rows f-1/f-2/f-3): FASTQ text in pinned host memory -> atr_trim_fastq_host -> trimmed FASTQ
    # text + the report's statistics in host memory. Same reads; reader, trimming and formatting on the GPU. ----
    fq = None
    if not args.no_fastq:
        from atropos_b200 import fastq as fastq_mod
        n_fq = min(n, args.fastq_reads)                        # the same per-GPU size at every N
        text_np = synth.fastq_text(reads_host.numpy()[:n_fq])
        text_host = torch.empty(text_np.size, dtype=torch.uint8, pin_memory=True)
        text_host.numpy()[:] = text_np
        del text_np
        fq_out = torch.empty(text_host.numel() + 1, dtype=torch.uint8, pin_memory=True)
        trimmer = fastq_mod.FastqTrimmer([adapter], times=1, max_len=L, device=rig.local)
        fq_steps = max(1, min(args.steps, 3))
        res = trimmer.trim(text_host.numpy(), out=fq_out.numpy())          # warm-up (allocations)
        ctx.launch_count(reset=True)
        box = {}

        def fq_step():
            box["res"] = trimmer.trim(text_host.numpy(), out=fq_out.numpy())
        fq_s = rig.time_host(fq_step, fq_steps, 0)
        fq_launches = ctx.launch_count() // fq_steps
        out_view, fq_stats, _ = box["res"]
        assert fq_stats.records == n_fq
        fq_link = rig.link_roof(int(text_host.numel()), int(out_view.size))
        fq = {"value": world * n_fq / fq_s / 1e6, "reads_per_gpu": n_fq, "unit": "M reads/s", "ms_per_step": fq_s * 1e3,
              "steps": fq_steps, "h2d_bytes_per_step": int(text_host.numel()), "d2h_bytes_per_step": int(out_view.size),
              "gpu_launches_per_step": int(fq_launches), "reads_with_adapters": int(fq_stats.with_adapters),
              "link_roof": dict(fq_link, M_reads_per_s=world * n_fq / fq_link["seconds_per_step"] / 1e6),
              "api": "atr_trim_fastq_host (fastq.FastqTrimmer.trim): FASTQ text -> trimmed FASTQ text + report statistics"}
        del text_host, fq_out, trimmer, res, box
    clocks = sampler.stop() if rank == 0 else None

    # parity spot check of the timed outputs (device leg vs host leg must agree bit for bit)
    a = out_dev.cpu().numpy().view(_abi.MATCH_DTYPE).reshape(-1)
    b = out_host.numpy().view(_abi.MATCH_DTYPE).reshape(-1)
    assert np.array_equal(a, b), "device-resident and host entry points disagree"
    hit_frac = float((a["status"] == _abi.ATR_ST_MATCH).mean())
    if rank != 0:
        return None

    dom_name, dom_ms = ("step", ms_per_step)
    if kernels_ms:
        dom_name = max(kernels_ms, key=kernels_ms.get)
        dom_ms = kernels_ms[dom_name]
    line = {
        "metric": "M reads/sec trimmed (150 bp SE, TruSeq 3' adapter, err 0.1)",
        "value": value, "unit": "M reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 packed integer DP (4-bit bases)", "data": "synthetic",
        "config": dict(CONFIG, reads_per_gpu=n, adapter_hit_fraction=round(hit_frac, 4), host_binding=rig.host_binding,
                       same_bytes="the first %d reads of shard 0 come from the CPU generator: what --impl reference times" % (prefix.shape[0] if prefix is not None else 0)),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "M reads/s", "h2d_bytes_per_step": int(n * L),      # fixed-length batch: the offsets are rebuilt on the device
                "d2h_bytes_per_step": int(16 * n), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "api": "atr_locate_batch_host (Adapter.match_to_batch)", "link_roof": link},
        "e2e_packed": e2e_packed,
        "e2e_fastq": fq,
        "strong_scaling": strong,
        "gpu_launches": int(launches),
        "roofline": roofline_block(rig, ALGO_BYTES_PER_READ, n, dom_name, dom_ms, kernels_ms, ms_per_step, {
            "note": "integer-ALU bound path: see DESIGN.md; equivalent full-matrix cell rate below",
            "gcups_equivalent": n * READ_LEN * len(ADAPTER) / (ms_per_step * 1e-3) / 1e9}),
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        pre = prefix if prefix is not None else reads_host.numpy()[:CPU_PREFIX]
        sample = min(pre.shape[0], 100_000 * cores)
        rate, secs, kind = cpu_reference_rate(pre, sample, cores)
        r1, s1, _ = cpu_reference_rate(pre, 300_000, 1)
        cli = {} if args.no_cli_baseline else reference_cli_rates(pre, cores, budget_s=60.0)
        line["cpu_baseline"] = {"value": rate, "unit": "M reads/s", "cores": cores, "kind": kind,
                                "sample": "first %d reads of the same batch, reference Aligner.locate (compiled Cython), "
                                          "one process per core, %.1f s; single core: %.3f M reads/s" % (sample, secs, r1),
                                "reference_cli": cli}
    return line


# ------------------------------------------------------------------------------------------------
# cfg 3 / cfg 5: InsertAligner.match_insert over pairs
# ------------------------------------------------------------------------------------------------
def oracle_check_pairs(r1, r2, L, rate, count, recs):
    from atropos_b200 import synth
    from atropos_b200.align import InsertAligner
    from oracle import oracle as orc
    oia = orc.OracleInsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, max_insert_mismatch_frac=rate, max_adapter_mismatch_frac=rate)
    bad = matched = 0
    for i in range(count):
        e = oia.match_insert(bytes(r1[i]).decode(), bytes(r2[i]).decode())
        g = InsertAligner.result_from_record(recs[i])
        if e is None:
            bad += g is not None
        else:
            matched += 1
            ok = g is not None and g[0] == e[0] and (g[1].fields() if g[1] else None) == e[1] and \
                (g[2].fields() if g[2] else None) == e[2]
            bad += not ok
    return {"units": count, "mismatches": int(bad), "oracle_matches": int(matched),
            "checker": "oracle.OracleInsertAligner (CPU restatement of InsertAligner.match_insert) on the first pairs of the timed batch"}


def run_pairs(rig, args, cfg, L, rate, pairs_per_gpu, chunks, label):
    import numpy as np
    torch = rig.torch
    from atropos_b200 import _abi, synth
    from atropos_b200.align import InsertAligner
    dev, world, rank = rig.dev, rig.world, rig.rank
    sub = 0.02 if cfg == 5 else 0.01
    ia = InsertAligner(synth.TRUSEQ_R1, synth.TRUSEQ_R2, max_insert_mismatch_frac=rate, max_adapter_mismatch_frac=rate,
                       device=rig.local)
    iset = ia._insertset(L)
    algo = 2 * ((L + 1) // 2) + 8 + 48
    steps = max(1, min(args.steps, 5))
    tot_ms, tot_pairs, launches, hit, first = 0.0, 0, 0, [], None
    n = pairs_per_gpu
    for c in range(chunks):
        r1, r2 = synth.synth_pe(n, L, seed=synth.seed_for(cfg, rank * 1000 + c), device=dev, sub=sub)
        c1, w1, l1, _ = rig.pack(r1, False)
        c2, w2, l2, _ = rig.pack(r2, False)
        iout = torch.empty((n, 48), dtype=torch.uint8, device=dev)
        ms, nl = rig.time_device(lambda: iset.match_insert_device(c1.data_ptr(), w1.data_ptr(), l1.data_ptr(), c2.data_ptr(),
                                                                   w2.data_ptr(), l2.data_ptr(), n, iout.data_ptr()),
                                 steps, 2 if c == 0 else 1)
        tot_ms += ms
        tot_pairs += n
        launches += nl // steps
        res = iout.cpu().numpy().view(_abi.INSERT_DTYPE).reshape(-1)
        hit.append(float((res["insert"]["status"] == _abi.ATR_ST_MATCH).mean()))
        if c == 0:
            first = (r1, r2, res)
        else:
            del r1, r2
        del c1, w1, l1, c2, w2, l2, iout
    r1, r2, res = first
    # e2e: the first chunk through the host entry point (ASCII pairs in pinned host memory -> 48-byte records in host memory)
    h1, h2 = rig.pinned_copy(r1), rig.pinned_copy(r2)
    offs = np.arange(n + 1, dtype=np.int64) * L
    out_host = torch.empty((n, 48), dtype=torch.uint8, pin_memory=True)
    ov = out_host.numpy().view(_abi.INSERT_DTYPE).reshape(-1)
    e2e_s = rig.time_host(lambda: iset.match_insert_host(h1.numpy().reshape(-1), offs, h2.numpy().reshape(-1), offs, out=ov), 2, 1)
    assert np.array_equal(ov, res), "device-resident and host entry points disagree (%s)" % label
    link = rig.link_roof(int(2 * n * L + 16 * (n + 1)), int(48 * n), steps=2)
    check = None
    if rank == 0 and not args.no_oracle_check:
        check = oracle_check_pairs(h1.numpy(), h2.numpy(), L, rate, min(n, args.check_units), res)
    value = world * tot_pairs / (tot_ms * 1e-3) / 1e6
    e2e_value = world * n / e2e_s / 1e6
    if rank != 0:
        return None
    ms_per_chunk = tot_ms / chunks
    kname = "k_insert_packed<true>" if int(rate * L) >= 18 else "k_insert_packed<false>"
    return {"workload": label, "value": value, "unit": "M pairs/s", "pairs_per_gpu": tot_pairs, "chunks": chunks,
            "pairs_per_chunk": n, "ms_per_chunk": ms_per_chunk, "gpu_launches_per_chunk": launches // chunks,
            "insert_match_fraction": sum(hit) / len(hit),
            "roofline": roofline_block(rig, algo, n, kname, ms_per_chunk, {kname: ms_per_chunk}, ms_per_chunk),
            "e2e": {"value": e2e_value, "unit": "M pairs/s", "h2d_bytes_per_step": int(2 * n * L + 16 * (n + 1)),
                    "d2h_bytes_per_step": int(48 * n), "ms_per_step": e2e_s * 1e3,
                    "api": "atr_match_insert_batch_host (InsertAligner.match_insert_batch)",
                    "link_roof": dict(link, M_pairs_per_s=world * n / link["seconds_per_step"] / 1e6)},
            "oracle_check": check}


# ------------------------------------------------------------------------------------------------
# cfg 4: panel of 8 adapters, best match per read
# ------------------------------------------------------------------------------------------------
def run_cfg4(rig, args):
    import numpy as np
    torch = rig.torch
    from atropos_b200 import _abi, adapters as ad_mod, synth
    from atropos_b200.modifiers import AdapterCutter
    dev, world, rank = rig.dev, rig.world, rig.rank
    L = READ_LEN
    n = args.panel_reads or (100_000_000 // max(world, 4))            # 25 M per GPU up to 4 GPUs: 100 M at N = 4
    reads = synth.synth_se(n, L, ADAPTER, seed=synth.seed_for(4, rank), device=dev)
    # 10 % of the reads carry a 5' construct: one of the two anchored 5' adapters at the read start
    g = torch.Generator(device=dev)
    g.manual_seed(synth.seed_for(4, rank) + 7)
    pick = torch.rand(n, generator=g, device=dev)
    for k, seq in enumerate([PANEL[3][0], PANEL[4][0]]):
        rows = torch.nonzero((pick >= 0.05 * k) & (pick < 0.05 * (k + 1))).squeeze(1)
        reads[rows, :len(seq)] = torch.tensor(list(seq.encode()), dtype=torch.uint8, device=dev)[None, :]
    cutter = AdapterCutter([ad_mod.Adapter(s, getattr(ad_mod, w), max_error_rate=ERROR_RATE, min_overlap=MIN_OVERLAP,
                                           device=rig.local) for s, w in PANEL], device=rig.local)
    pset = cutter._adapterset()
    codes, woff, lens, _ = rig.pack(reads, True)
    out_dev = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    steps = max(1, min(args.steps, 5))
    ms, launches = rig.time_device(lambda: pset.locate_device(codes.data_ptr(), woff.data_ptr(), lens.data_ptr(), n, out_dev.data_ptr()),
                                   steps, 2)
    res = out_dev.cpu().numpy().view(_abi.MATCH_DTYPE).reshape(-1)
    del codes, woff, lens
    reads_host = rig.pinned_copy(reads)
    del reads
    offs = np.arange(n + 1, dtype=np.int64) * L
    out_host = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
    ov = out_host.numpy().view(_abi.MATCH_DTYPE).reshape(-1)
    e2e_s = rig.time_host(lambda: pset.locate_host(reads_host.numpy().reshape(-1), offs, fold_case=True, out=ov), 2, 1)
    assert np.array_equal(ov, res), "device-resident and host entry points disagree (cfg4)"
    link = rig.link_roof(int(n * L), int(16 * n), steps=2)
    check = None
    if rank == 0 and not args.no_oracle_check:
        from oracle import oracle as orc
        oads = [orc.OracleAdapter(s, getattr(orc, w), ERROR_RATE, MIN_OVERLAP) for s, w in PANEL]
        cnt = min(n, args.check_units)
        bad = hits = 0
        rh = reads_host.numpy()
        for i in range(cnt):
            seq = bytes(rh[i]).decode()
            best, bi = None, -1
            for ai, oa in enumerate(oads):                      # AdapterCutter._best_match (modifiers.py:107-122)
                m = oa.match_to(seq)
                if m is not None and (best is None or m[4] > best[4]):
                    best, bi = m, ai
            gr = res[i]
            if best is None:
                bad += int(gr["status"]) != _abi.ATR_ST_NONE
            else:
                hits += 1
                got = tuple(int(gr[k]) for k in ("astart", "astop", "rstart", "rstop", "matches", "errors"))
                bad += not (int(gr["status"]) == _abi.ATR_ST_MATCH and got == tuple(best[:6]) and int(gr["adapter"]) == bi)
        check = {"units": cnt, "mismatches": int(bad), "oracle_matches": int(hits),
                 "checker": "oracle.OracleAdapter.match_to for all 8 adapters + the reference's best-match rule, first reads of the timed batch"}
    value = world * n / (ms * 1e-3) / 1e6
    if rank != 0:
        return None
    winners = np.bincount(res["adapter"][res["status"] == _abi.ATR_ST_MATCH].astype(np.int64), minlength=len(PANEL))
    return {"workload": "cfg4: SE 150 bp, 8-adapter panel (3x 3', 2x anchored 5', 1x 5', linked adapters as their components), "
                        "best match per read on the GPU, %d M reads per GPU" % (n // 1_000_000),
            "value": value, "unit": "M reads/s", "reads_per_gpu": n, "reads_total": world * n, "ms_per_step": ms,
            "gpu_launches_per_step": launches // steps, "match_fraction": float((res["status"] == _abi.ATR_ST_MATCH).mean()),
            "winning_adapter_histogram": [int(x) for x in winners],
            "roofline": roofline_block(rig, ALGO_BYTES_PER_READ, n, "panel step (all kernels of the 8 adapters)", ms, None, ms),
            "e2e": {"value": world * n / e2e_s / 1e6, "unit": "M reads/s", "h2d_bytes_per_step": int(n * L),
                    "d2h_bytes_per_step": int(16 * n), "ms_per_step": e2e_s * 1e3,
                    "api": "atr_locate_batch_host (AdapterCutter.best_match_batch)",
                    "link_roof": dict(link, M_reads_per_s=world * n / link["seconds_per_step"] / 1e6)},
            "oracle_check": check}


def run_ours(args):
    rig = Rig(args)
    line = run_cfg2(rig, args)
    rig.torch.cuda.empty_cache()
    configs = {}
    if not args.no_configs:
        t0 = time.perf_counter()
        configs["cfg3"] = run_pairs(rig, args, 3, 150, 0.1, args.pairs or 10_000_000, 1,
                                    "cfg3: PE 2x150, --aligner insert (TruSeq R1/R2), err 0.1, 10 M pairs per GPU")
        rig.torch.cuda.empty_cache()
        configs["cfg4"] = run_cfg4(rig, args)
        rig.torch.cuda.empty_cache()
        configs["cfg5"] = run_pairs(rig, args, 5, 300, 0.15, 4_000_000, args.cfg5_chunks,
                                    "cfg5: PE 2x300, err 0.15, insert aligner, streamed in 4 M-pair chunks generated on the fly "
                                    "(--cfg5-chunks 31 per GPU at 8 GPUs = the 1 B-pair job)")
        if rig.rank == 0:
            configs["seconds"] = time.perf_counter() - t0
    if rig.rank == 0:
        line["configs"] = configs or None
        print(json.dumps(line))
    if rig.world > 1:
        rig.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the 10 M of BASELINE config 2)")
    ap.add_argument("--pairs", type=int, default=0, help="cfg3 pairs per GPU (default 10 M)")
    ap.add_argument("--panel-reads", type=int, default=0, help="cfg4 reads per GPU (default 100 M / max(N, 4))")
    ap.add_argument("--cfg5-chunks", type=int, default=2, help="cfg5: 4 M-pair chunks per GPU (31 at 8 GPUs = 1 B pairs)")
    ap.add_argument("--fastq-reads", type=int, default=4_000_000, help="reads per GPU of the FASTQ-text leg")
    ap.add_argument("--check-units", type=int, default=100_000, help="reads / pairs per config checked against the oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli-baseline", action="store_true", help="skip the `atropos trim -T N` runs of the reference package")
    ap.add_argument("--no-fastq", action="store_true", help="skip the FASTQ-text leg (e2e_fastq)")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg4 / cfg5 blocks")
    ap.add_argument("--no-oracle-check", action="store_true")
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
