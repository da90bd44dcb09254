#!/usr/bin/env python
"""Instruction and stall-sample share per code region of an ncu report (source page).

usage: tools/ncu_regions.py report.ncu-rep [--json out.json]

A region is a device function (any `__device__ ...` definition at column 0, whatever its
inlining macro, in brs_kernels.cuh and brs_wide.cuh) or a kernel, split further at the
phase banners (`// ---------------- NAME ---...`) inside the long ones.  Lines are
attributed to the innermost region that starts at or before them in the same file.
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = {n: os.path.join(ROOT, "aitk.robots_b200", "csrc", n) for n in ("brs_kernels.cuh", "brs_wide.cuh", "brs_solo.cuh", "brs_multi.cuh")}


def region_marks(path):
    marks, fn = [], "?"
    lines = open(path).read().splitlines()
    for i, l in enumerate(lines, 1):
        if l.startswith("__device__"):
            m = re.search(r"(\w+)\s*\(", l)
            if m:
                fn = m.group(1)
                marks.append((i, fn))
        elif l.startswith("__global__"):
            for k in range(i - 1, min(i + 3, len(lines))):
                m = re.search(r"(brs_\w+_kernel)", lines[k])
                if m:
                    fn = m.group(1)
                    marks.append((i, fn))
                    break
        else:
            m = re.match(r"\s*// -{8,} ([A-Za-z0-9 +,:/]+?) -{3,}", l)
            if m:
                marks.append((i, fn + ":" + m.group(1).strip()))
            elif "5a. group path" in l:
                marks.append((i, fn + ":SENSE group path"))
            elif "5b. direct path" in l:
                marks.append((i, fn + ":SENSE direct path"))
    return marks


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, fname, data = None, "", {}
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) >= len(hdr) - 2 and r[0].isdigit():
            data.setdefault((fname, int(r[0])), r)
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0  # noqa: E731
    marks = {n: region_marks(p) for n, p in SRC.items() if os.path.exists(p)}
    tot_i = sum(num(r[ci]) for r in data.values())
    tot_s = sum(num(r[cs]) for r in data.values())
    agg = {}
    for (f, ln), r in data.items():
        name = "other:" + f
        if f in marks:
            name = "?"
            for i, n in marks[f]:
                if i <= ln:
                    name = n
                else:
                    break
        a = agg.setdefault(name, {"inst": 0, "samples": 0, "stalls": {}})
        a["inst"] += num(r[ci])
        a["samples"] += num(r[cs])
        for i, n in stall_cols:
            v = num(r[i])
            if v:
                a["stalls"][n] = a["stalls"].get(n, 0) + v
    print("total warp-inst %d samples %d" % (tot_i, tot_s))
    res = []
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"]):
        top = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
        print("%-44s inst %5.1f%%  samples %5.1f%%  %s" % (n, 100.0 * a["inst"] / max(tot_i, 1), 100.0 * a["samples"] / max(tot_s, 1),
                                                          ",".join("%s:%d" % kv for kv in top)))
        res.append({"region": n, "inst_pct": 100.0 * a["inst"] / max(tot_i, 1), "samples_pct": 100.0 * a["samples"] / max(tot_s, 1)})
    if "--json" in sys.argv:
        json.dump({"total_warp_inst": tot_i, "samples": tot_s, "regions": res}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
