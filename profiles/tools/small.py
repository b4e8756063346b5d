"""Planner-sized batches only (device-resident, L2 flushed, median of 40): A/B of knobs or builds.
usage: [RC_LIB=...] [RC_SPLIT_MAX=..] python profiles/tools/small.py [tag]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
s = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ctx = rc.Context(device=0)
out = {"tag": tag}
for name, n, hard, seed in (("far2k", 2000, False, 510), ("hard64", 64, True, 508), ("hard512", 512, True, 508),
                            ("hard2k", 2000, True, 508), ("hard10k", 10000, True, 508), ("hard30k", 30000, True, 508),
                            ("hard100k", 100000, True, 508)):
    confs = rc.createAlphaBotConfigurations(n, hard, seed)
    if name == "far2k":
        confs[:, :3] += 5000.0
    d = torch.from_numpy(confs).cuda()
    v = torch.zeros(n, dtype=torch.uint8, device="cuda")
    m = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.check(d, n, v, m, rc.RC_MODE_CULLED, s)
    torch.cuda.synchronize()
    ts = []
    for _ in range(40):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); ctx.check(d, n, v, m, rc.RC_MODE_CULLED, s); e1.record(s)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    out[name] = round(ts[len(ts) // 2] * 1e3, 1)
    out[name + "_valid"] = int(v.sum().item())
print(json.dumps(out), flush=True)
