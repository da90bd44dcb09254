#!/usr/bin/env python
"""Regenerates profiles/<round>/README.md from the committed bench lines (x_bench_lines.json)."""
import glob
import json
import os
import sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", rnd)
out = ["# profiles/%s -- measured on one B200 (gpurun), CUDA-event timing, ncu summaries\n" % rnd,
       "Files: `*_bench_lines.json` = the JSON lines `bench.py` printed; `*_ncu_summary_*.txt` = key metrics of one",
       "`ncu --set full --clock-control none` capture of `brs_step_kernel` + instruction share per code region;",
       "`*_ncu_lines_*.txt` = top source lines by stall samples; `*_launches_*.csv` = `gpu__time_duration` per launch.",
       "Prefix `a_` = per-world CTA kernel the round started from, later prefixes = the tile kernel as it evolved.\n"]
for f in sorted(glob.glob(os.path.join(d, "*_bench_lines.json"))):
    lines = json.load(open(f))
    out.append("## %s\n" % os.path.basename(f))
    out.append("| workload | worlds/GPU | ms/step | world-steps/s | ray-seg tests/s | HBM frac (of measured copy peak) | FP32 frac (13 flop/test, of measured FFMA peak) | e2e world-steps/s |")
    out.append("|---|---|---|---|---|---|---|---|")
    for l in lines:
        r = l["roofline"]
        hb = r["hbm"]["frac"] if "hbm" in r else r["frac"]
        fp = r["fp32"]["frac"] if "frac" in r["fp32"] else r["fp32"].get("frac_measured") or 0.0
        out.append("| %s | %d | %.4f | %.3g | %.3g | %.3f | %.3f | %.3g |" % (
            l["config"]["workload"].split(":")[0], l["config"]["worlds_per_gpu"], l["ms_per_step"], l["value"],
            l["ray_segment_tests_per_sec"], hb, fp, l["e2e"]["value"]))
    out.append("")
open(os.path.join(d, "README.md"), "w").write("\n".join(out))
print("\n".join(out))
