#!/usr/bin/env python
"""Instruction share of each phase of check_culled_kernel: joins the SASS page of an ncu report
(ncu -i X.ncu-rep --page source --csv --print-source sass) with nvdisasm -g line info of the cubin of
the SAME build, and maps source lines to phases through the marker comments in rc_culled.cuh and the
function headers in rc_device.cuh / rc_kernels.cu.
usage: phases.py <sass.csv> <kernels.dis> <mangled-kernel-substring> <csrc dir>"""
import csv
import os
import re
import sys
from collections import defaultdict

sass_csv, dis, kern, csrc = sys.argv[1:5]


def markers(path, pats):
    out = []
    for n, l in enumerate(open(path), 1):
        for pat, name in pats:
            if re.search(pat, l):
                out.append((n, name))
    return sorted(out)


cul = markers(os.path.join(csrc, "rc_culled.cuh"), [
    (r"bool boxes_overlap\(", "B node round: box test"), (r"bool narrow_exhaustive\(", "overflow fallback"),
    (r"struct ClassifyArgs", "classify"), (r"^__global__.*check_culled_kernel", "prologue / table staging"),
    (r"auto obs_node = ", "obstacle table reads (B, C, D rounds)"), (r"if \(threadIdx.x == 0\) \{\s*$", "prologue / table staging"),
    (r"Guided self-scheduling", "work fetch / serve loop"), (r"// A: exact broad-phase flag", "A exact broad phase"),
    (r"the descent starts from a cut", "seeds"), (r"// B \+ C \+ D", "loop control / hand-over"),
    (r"// ---- D: finish", "D finishing round"), (r"// ---- C: up to eight leaf pairs", "C leaf round"),
    (r"// ---- B: pop up to 32 node pairs", "B node round: queue code"), (r"if \(overflow\) \{", "publish / report"),
    (r"static size_t culled_smem_bytes", "host")])
dev = markers(os.path.join(csrc, "rc_device.cuh"), [
    (r"Mat34 make_matrix\(", "matrix"), (r"void xform\(", "vertex transform (A, C, D)"), (r"bool broad_valid\(", "A exact broad phase"),
    (r"Plane plane_of\(", "plane of robot triangle (C, D)"), (r"float sdist\(", "signed distances (C, D)"),
    (r"bool straddles\(", "plane-side test (C)"), (r"bool close0\(", "D finishing round (pair_tail)"),
    (r"uint32_t smem_u32\(", "prologue / table staging")])
ker = markers(os.path.join(csrc, "rc_kernels.cu"), [
    (r"float warp_min\(", "A exact broad phase"), (r"Mat34 load_matrix\(", "matrix"), (r"void warp_aabb\(", "A exact broad phase"),
    (r"void publish\(", "publish / report"), (r"bool exhaustive_sweep\(", "exhaustive")])
tables = {"rc_culled.cuh": cul, "rc_device.cuh": dev, "rc_kernels.cu": ker}


def phase(f, ln):
    t = tables.get(f)
    if t is None:
        return "CUDA intrinsics headers (shuffle / vote / atomics)"
    name = "?"
    for n, nm in t:
        if ln >= n:
            name = nm
    return name


line_of, cur, inside = {}, None, False
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
ia, ii, it, isamp = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
agg, base = defaultdict(lambda: [0, 0, 0]), None
for r in rows[2:]:
    try:
        addr = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = addr
    c = line_of.get(addr - base)
    a = agg[phase(*c) if c else "unknown"]
    a[0] += int(r[ii]); a[1] += int(r[it]); a[2] += int(r[isamp])
tot = sum(a[0] for a in agg.values()) or 1
ts = sum(a[2] for a in agg.values()) or 1
print("| phase | warp instructions | share | lanes / instruction | pc samples |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"| {k} | {a[0]:,} | {100 * a[0] / tot:.1f} % | {a[1] / max(a[0], 1):.1f} | {100 * a[2] / ts:.1f} % |")
