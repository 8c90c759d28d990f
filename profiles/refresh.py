#!/usr/bin/env python
"""Turn the .ncu-rep captures in gpurun_out/ (scratch) into the tracked summaries under profiles/.

    python profiles/refresh.py <tag>      # e.g. r1f: reads gpurun_out/prof_<kernel>_<tag>.ncu-rep, launches_<tag>.csv,
                                          # bench_<tag>.log and writes profiles/<round>_*.csv / traffic.json
"""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active"]
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def main(tag, reads=10_000_000):
    traffic = {}
    summary = []
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*_%s.ncu-rep" % tag))):
        kernel = os.path.basename(rep)[len("prof_"):-len("_%s.ncu-rep" % tag)]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        dst = os.path.join(PROF, "r1_%s_full_raw.csv" % kernel)
        open(dst, "w").write(raw)
        rows = list(csv.reader(raw.splitlines()))
        m = dict(zip(rows[0], zip(rows[1], rows[2])))
        tot = sum(float(m[k][1]) * UNIT[m[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[kernel] = {"dram_bytes_per_launch": tot, "dram_bytes_per_read": tot / reads,
                           "gpu_time": " ".join(reversed(m["gpu__time_duration.sum"]))}
        summary.append((kernel, {k: " ".join(reversed(m[k])) for k in KEYS if k in m}))
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        lines = src.splitlines()
        with open(os.path.join(PROF, "r1_%s_source_hot.csv" % kernel), "w") as fh:
            for ln in lines:
                fh.write(",".join(next(csv.reader([ln]))[:8]) + "\n") if ln.startswith('"0x') or ln.startswith('"Address') else None
    # merge into what is tracked already: a refresh may cover only the kernels that changed
    tpath, spath = os.path.join(PROF, "traffic.json"), os.path.join(PROF, "r1_kernel_summary.json")
    old_t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    old_s = json.load(open(spath)) if os.path.exists(spath) else {}
    old_t.update(traffic)
    old_s.update(dict(summary))
    json.dump(old_t, open(tpath, "w"), indent=1)
    json.dump(old_s, open(spath, "w"), indent=1)
    for name, dst in (("launches_%s.csv" % tag, "r1_launches_fastpath.csv"), ("bench_%s.log" % tag, "r1_bench_line.json")):
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(PROF, dst))
    print(json.dumps(dict(summary), indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
