#!/usr/bin/env python
"""Join an ncu capture's per-instruction counters with source lines.

    python tools/ncu_by_line.py <rep.ncu-rep> <mangled-or-substring kernel name> [--so atropos_b200/libatropos_b200.so] [--top 40]

ncu's `--page source --csv` lists the SASS of the profiled kernel in address order with executed-instruction counts;
`nvdisasm -g` on the cubin of the same build lists the same SASS with `//## File ..., line N` markers (build with
-lineinfo). Instruction k of one is instruction k of the other. Prints warp-instructions executed per source line
(inlined callee lines keep their own file:line), with the average number of active lanes.
"""
import argparse
import csv
import os
import re
import subprocess
import sys
import tempfile


def ncu_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = []
    for r in csv.reader(out.splitlines()):
        if r and r[0].startswith("0x"):
            nums = r[-1:]
            rows.append(r)
    return rows


def line_map(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    lines = []
    for cb in cubins:
        dis = subprocess.run(["nvdisasm", "-g", cb], capture_output=True, text=True).stdout.splitlines()
        inside = False
        cur = ("?", 0)
        for ln in dis:
            if ln.startswith("//--------------------- .text."):
                inside = kernel in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln) and "/*" in ln:
                body = ln.split("*/", 1)[1].strip()
                if body and not body.startswith("."):
                    lines.append((cur, body))
        if lines:
            break
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--so", default="atropos_b200/libatropos_b200.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--src-root", default="atropos_b200/csrc")
    a = ap.parse_args()
    rows = ncu_rows(a.rep)
    lm = line_map(a.so, a.kernel)
    if len(rows) != len(lm):
        print("warning: %d instructions in the capture, %d in the disassembly (different build?)" % (len(rows), len(lm)), file=sys.stderr)
    # columns: Address, Source, stall all, stall not issued, samples, inst executed, thread inst executed, ...
    agg = {}
    tot = 0
    for r, (loc, body) in zip(rows, lm):
        # the Source column may contain commas: numeric columns are taken relative to the end-stable layout
        src_end = 2
        while src_end < len(r) and not re.fullmatch(r"\d+", r[src_end]):
            src_end += 1
        nums = r[src_end:]
        inst, tinst, samples = int(nums[3]), int(nums[4]), int(nums[2])
        e = agg.setdefault(loc, [0, 0, 0, 0])
        e[0] += inst; e[1] += tinst; e[2] += samples; e[3] += 1
        tot += inst
    print("total warp instructions executed: %d over %d SASS instructions" % (tot, len(rows)))
    cache = {}
    for loc, e in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        f, n = loc
        text = ""
        p = os.path.join(a.src_root, f)
        if os.path.exists(p):
            if p not in cache:
                cache[p] = open(p).read().splitlines()
            if 0 < n <= len(cache[p]):
                text = cache[p][n - 1].strip()[:90]
        print("%5.1f%%  %10d inst  lanes %4.1f  samples %6d  sass %3d  %s:%d  %s" % (
            100.0 * e[0] / tot, e[0], e[1] / max(e[0], 1), e[2], e[3], f, n, text))


if __name__ == "__main__":
    main()
