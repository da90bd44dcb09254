#!/usr/bin/env python
"""Copy what one gpurun call measured (gpurun_out/<tag>/) into profiles/r2/ as text summaries:
bench lines, per-capture ncu key metrics + instruction share per code region, hot lines, and
the pipe-utilisation / DRAM-traffic numbers bench.py quotes (pipes.json, traffic_<config>.json).

usage: tools/collect_profiles.py <tag> <prefix> [capture=config[:steps_per_launch] ...]   e.g.  r2g a c4_solo=config4
       (a capture of a brs_steps launch names how many steps the launch ran: config2_steps=config2:10)
"""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, prefix = sys.argv[1], sys.argv[2]
capmap = dict(a.split("=") for a in sys.argv[3:])
src, dst = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles", "r2")
os.makedirs(dst, exist_ok=True)
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()

lines = []
for f in sorted(glob.glob(os.path.join(src, "b_*.json")) + glob.glob(os.path.join(src, "bench_*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        d["_run"] = os.path.basename(f)[:-5]
        lines.append(d)
    except Exception:
        pass
if lines:
    json.dump(lines, open(os.path.join(dst, "%s_bench_lines.json" % prefix), "w"), indent=1)

pipes_path = os.path.join(dst, "pipes.json")
pipes = json.load(open(pipes_path)) if os.path.exists(pipes_path) else {}
for rep in sorted(glob.glob(os.path.join(src, "prof_*.ncu-rep"))):
    name = os.path.basename(rep)[5:-8]
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    regs = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep], capture_output=True, text=True).stdout
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "40"], capture_output=True, text=True).stdout
    open(os.path.join(dst, "%s_ncu_summary_%s.txt" % (prefix, name)), "w").write(
        "# ncu --set full --clock-control none, one launch; gpurun call %s, source at or after commit %s\n" % (tag, head)
        + summ + "\n# instruction and stall-sample share per code region (tools/ncu_regions.py)\n" + regs)
    open(os.path.join(dst, "%s_ncu_lines_%s.txt" % (prefix, name)), "w").write(hot)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    m = dict(zip(rows[0], rows[2]))
    num = lambda k: float(m[k].replace(",", "")) if k in m and m[k] not in ("", "n/a") else None  # noqa: E731
    cfg = capmap.get(name)
    spl = 1
    if cfg and ":" in cfg:
        cfg, spl = cfg.split(":")[0], int(cfg.split(":")[1])
    if cfg:
        pipes[cfg] = {
            "kernel": m.get("Kernel Name"), "source": "profiles/r2/%s_ncu_summary_%s.txt" % (prefix, name),
            "issue_slots_busy_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": num("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active"),
            "alu_pipe_pct": num("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
            "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"),
            "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active"),
            "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "shared_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "shared_bank_conflicts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            "warp_instructions": num("smsp__inst_executed.sum"),
            "gpu_time_us_under_ncu": num("gpu__time_duration.sum"), "steps_per_launch": spl,
        }
        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            u = dict(zip(rows[0], rows[1]))
            tot = rd * unit.get(u.get("dram__bytes_read.sum"), 1.0) + wr * unit.get(u.get("dram__bytes_write.sum"), 1.0)
            json.dump({"dram_bytes_per_launch": tot, "steps_per_launch": spl, "dram_bytes_per_step": tot / spl, "source": "ncu --set full, %s_ncu_summary_%s.txt (gpurun call %s, commit %s)"
                       % (prefix, name, tag, head)}, open(os.path.join(dst, "traffic_%s.json" % cfg), "w"), indent=1)
json.dump(pipes, open(pipes_path, "w"), indent=1)
for f in glob.glob(os.path.join(src, "launches_*.csv")):
    keep = [l for l in open(f) if l.startswith('"') and ("gpu__time_duration" in l or l.startswith('"ID"'))]
    open(os.path.join(dst, "%s_%s" % (prefix, os.path.basename(f))), "w").writelines(keep)
print("collected", tag, "->", dst, ":", len(lines), "bench lines")
