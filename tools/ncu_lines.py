#!/usr/bin/env python
"""Per-source-line summary of an ncu report (instructions executed, stall samples).
usage: tools/ncu_lines.py report.ncu-rep [top_n] [kernel-instance]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, fname, data, seen_fn = None, "", [], set()
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) >= len(hdr) - 2 and r[0] not in ("", "Line No"):
        data.append((fname, r))
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0  # noqa: E731
# several kernel instances repeat the same lines: keep the first instance of each (file, line)
uniq = {}
for f, r in data:
    uniq.setdefault((f, r[0]), r)
data = [(k[0], r) for k, r in uniq.items()]
tot_i = sum(num(r[ci]) for _, r in data)
tot_s = sum(num(r[cs]) for _, r in data)
print("total warp-inst %d, samples %d" % (tot_i, tot_s))
agg = sorted(((sum(num(r[i]) for _, r in data), n) for i, n in stalls), reverse=True)
print("stall totals:", [(n, v) for v, n in agg[:9] if v])
data.sort(key=lambda fr: -num(fr[1][cs]))
for f, r in data[:top]:
    st = sorted(((num(r[i]), n) for i, n in stalls), reverse=True)[:2]
    print("%-16s %4s inst %5.1f%% smp %5.1f%% %-30s| %s" % (f[:16], r[0], 100.0 * num(r[ci]) / max(tot_i, 1),
          100.0 * num(r[cs]) / max(tot_s, 1), ",".join("%s:%d" % (n, v) for v, n in st if v), r[1].strip()[:80]))
