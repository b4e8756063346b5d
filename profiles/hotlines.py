#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line
info of the cubin, and print instruction counts per CUDA source line.
usage: hotlines.py <sass.csv> <kernels.dis> <mangled-kernel-substring> [top]"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. nvdisasm: walk the kernel's section, track the current (file, line) per instruction offset
line_of = {}
cur = None
inside = False
for l in open(dis):
    if l.startswith("//---------------------"):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    try:
        addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))
    vals = [int(float(r[ii] or 0)), int(float(r[it] or 0)), int(float(r[isamp] or 0))]
    for k in range(3):
        agg[key][k] += vals[k]
        tot[k] += vals[k]
print(f"total warp-inst {tot[0]:,}  thread-inst {tot[1]:,}  samples {tot[2]:,}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]:16s}:{key[1]:4d}  inst {v[0]:12,} ({100*v[0]/tot[0]:5.1f}%)  lanes {v[1]/max(v[0],1):5.1f}  samples {100*v[2]/max(tot[2],1):5.1f}%")
