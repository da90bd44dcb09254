#!/usr/bin/env python
"""Instruction share per source-line range. usage: tools/ncu_regions.py report.ncu-rep"""
import csv, subprocess, sys, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr=None; fname=""; data={}
for r in rows:
    if r and r[0]=="File Path": fname=r[1].split("/")[-1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)>=len(hdr)-2 and r[0].isdigit():
        key=(fname,int(r[0]))
        if key not in data: data[key]=r
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
num=lambda v:int(v) if v.strip().lstrip("-").isdigit() else 0
src=open("/root/repo/aitk.robots_b200/csrc/brs_kernels.cuh").read().splitlines()
# region markers: function definitions and phase comments
marks=[]
for i,l in enumerate(src,1):
    m=re.match(r"\s*// -+ phase (\d[^-]*)-",l) or re.match(r"^(?:template.*\n)?__device__ __forceinline__ \S+ (\w+)\(",l) or re.match(r"^__device__ __forceinline__ .*? (\w+)\(",l)
    if m: marks.append((i,m.group(1).strip()))
    elif "5a. group path" in l: marks.append((i,"5a group path"))
    elif "5b. direct path" in l: marks.append((i,"5b direct path"))
    elif "__global__" in l and "brs_step" in l: marks.append((i,"kernel prologue"))
    elif "for (int tile = sched" in l: marks.append((i,"tile loop head"))
marks.sort()
tot_i=sum(num(r[ci]) for (f,_),r in data.items()); tot_s=sum(num(r[cs]) for (f,_),r in data.items())
agg={}
for (f,ln),r in data.items():
    name="other:"+f
    if f=="brs_kernels.cuh":
        name="?"
        for i,n in marks:
            if i<=ln: name=n
            else: break
    a=agg.setdefault(name,[0,0]); a[0]+=num(r[ci]); a[1]+=num(r[cs])
print("total warp-inst %d samples %d"%(tot_i,tot_s))
for n,(i,s_) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print("%-28s inst %5.1f%%  samples %5.1f%%"%(n,100.0*i/tot_i,100.0*s_/max(tot_s,1)))
