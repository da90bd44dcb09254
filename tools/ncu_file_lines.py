#!/usr/bin/env python
"""Per-line executed warp-instructions (per world) of one source file in an ncu report, in
source order.  usage: tools/ncu_file_lines.py report.ncu-rep file.cuh first_line last_line n_worlds"""
import csv
import subprocess
import sys

rep, fn, a, b, nworlds = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
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
tot = 0
for (f, ln), r in sorted(data.items()):
    if f == fn and a <= ln <= b:
        n = int(r[ci]) if r[ci].isdigit() else 0
        tot += n
        if n:
            print("%4d %7.1f  smp %5s | %s" % (ln, n / nworlds, r[cs], r[1].strip()[:100]))
print("total per world: %.1f" % (tot / nworlds))
