"""Opcode histogram per kernel of librobocheck.so (cuobjdump -sass), with the mnemonics that prove the
Blackwell-native paths called out: UBLKCP (cp.async.bulk, TMA bulk copy), SYNCS (mbarrier), and the
absence of tensor-core opcodes (the path is FP32 geometry).  usage: sass_hist.py [lib.so] > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__)))), "robo-check_b200", "librobocheck.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# cuobjdump -sass {os.path.basename(lib)}: cubin architectures {arch}\n")
kern, hist = None, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["UBLKCP", "SYNCS", "UTMALDG", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMS", "ATOMG", "RED", "FFMA", "FMUL", "FADD",
       "FMNMX", "FMNMX3", "FSETP", "VOTE", "SHFL", "POPC", "MUFU", "BAR", "ACQBULK", "CCTL", "MEMBAR", "NANOSLEEP",
       "HMMA", "UTCHMMA", "LDTM"]
names = subprocess.run(["cu++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
for k, name in zip(hist, names):
    h = hist[k]
    short = name.split(">(")[0] + ">" if ">(" in name else re.sub(r"\(.*", "", name)
    total = sum(h.values())
    print(f"## {short}\n{total} instructions; " + ", ".join(f"{op} {h[op]}" for op in KEY if h[op]))
    rest = [(op, c) for op, c in h.most_common(12)]
    print("   top: " + ", ".join(f"{op} {c}" for op, c in rest) + "\n")
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("## whole library\n" + ", ".join(f"{op} {tot[op]}" for op in KEY))
