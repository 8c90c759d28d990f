#!/usr/bin/env python
"""Turn what tools/capture_profiles.sh <tag> left in gpurun_out/ (scratch) into the tracked summaries under profiles/.

    python profiles/refresh.py <tag> [round-prefix, default r2]

Writes  profiles/<round>_<kernel>_full_raw.csv      the ncu --set full raw page of one launch
        profiles/<round>_<kernel>_by_line.txt       warp instructions per source line of the same capture (tools/ncu_by_line.py)
        profiles/<round>_kernel_summary.json        the handful of metrics the README quotes, per kernel
        profiles/<round>_launches.csv               ncu launch list (gpu__time_duration) of the bench command
        profiles/<round>_bench_line.json            the un-profiled bench line of the same build
        profiles/traffic.json                       DRAM bytes per launch / per read of each kernel (bench.py's roofline.traffic)
"""
import csv
import glob
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}
UNITS = {"k_insert_packed_2x150": 10_000_000, "k_insert_packed_2x300": 4_000_000}     # pairs per launch; everything else: 10 M reads


def main(tag, rnd="r2"):
    traffic, summary = {}, {}
    for raw_path in sorted(glob.glob(os.path.join(OUT, "prof_*_%s_raw.csv" % tag))):
        kernel = os.path.basename(raw_path)[len("prof_"):-len("_%s_raw.csv" % tag)]
        shutil.copy(raw_path, os.path.join(PROF, "%s_%s_full_raw.csv" % (rnd, kernel)))
        bl = raw_path.replace("_raw.csv", "_by_line.txt")
        if os.path.exists(bl):
            shutil.copy(bl, os.path.join(PROF, "%s_%s_by_line.txt" % (rnd, kernel)))
        rows = list(csv.reader(open(raw_path)))
        m = dict(zip(rows[0], zip(rows[1], rows[2])))
        tot = sum(float(m[k][1]) * UNIT[m[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        units = UNITS.get(kernel, 10_000_000)
        traffic[kernel] = {"dram_bytes_per_launch": tot, "dram_bytes_per_read": tot / units, "units_per_launch": units,
                           "gpu_time": " ".join(reversed(m["gpu__time_duration.sum"])),
                           "capture": "%s_%s_full_raw.csv (ncu --set full --clock-control none, one launch, tag %s)" % (rnd, kernel, tag)}
        summary[kernel] = {k: " ".join(reversed(m[k])) for k in KEYS if k in m}
    tpath = os.path.join(PROF, "traffic.json")
    old_t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    old_t.update(traffic)
    json.dump(old_t, open(tpath, "w"), indent=1)
    json.dump(summary, open(os.path.join(PROF, "%s_kernel_summary.json" % rnd), "w"), indent=1)
    for name, dst in (("launches_%s.csv" % tag, "%s_launches.csv" % rnd), ("bench_%s.log" % tag, "%s_bench_line.json" % rnd),
                      ("panel_%s.log" % tag, "%s_panel_per_adapter.txt" % rnd), ("k2_150_%s.log" % tag, "%s_k2_2x150.txt" % rnd),
                      ("k2_300_%s.log" % tag, "%s_k2_2x300.txt" % rnd)):
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(PROF, dst))
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2a", sys.argv[2] if len(sys.argv) > 2 else "r2")
