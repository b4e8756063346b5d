"""Quick device-resident timings of the fused check on the workloads that matter, for A/B runs of two
builds on ONE box:   RC_LIB=ab/librobocheck_x.so python profiles/tools/quick.py [tag] [--c5]
Prints one line per workload: median and best of the event-timed launches (L2 flushed before each)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.environ.get("RC_LIB", "default")
s = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(ctx, confs, reps, mode=rc.RC_MODE_CULLED):
    n = len(confs)
    d = torch.from_numpy(confs).cuda()
    v = torch.zeros(n, dtype=torch.uint8, device="cuda")
    m = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.check(d, n, v, m, mode, s)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); ctx.check(d, n, v, m, mode, s); e1.record(s)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0], int(v.sum().item())


ctx = rc.Context(device=0)
rows = []
far = rc.createAlphaBotConfigurations(2_000, False, 510)
far[:, :3] += 5000.0   # every configuration cleared by step 1: the fixed cost of the two launches
med, best, nv = timeit(ctx, far, 40)
rows.append(("far2k_fixed", 2_000, med, best, nv))
for name, n, hard, seed, reps in (("easy2k_floor", 2_000, False, 510, 40), ("hard64", 64, True, 508, 40),
                                  ("hard100k", 100_000, True, 508, 40), ("hard10k", 10_000, True, 508, 40),
                                  ("hard2k", 2_000, True, 508, 40), ("hard1M", 1_000_000, True, 512, 8),
                                  ("easy1M", 1_000_000, False, 510, 20)):
    med, best, nv = timeit(ctx, rc.createAlphaBotConfigurations(n, hard, seed), reps)
    rows.append((name, n, med, best, nv))
if "--c5" in sys.argv:
    rob, obs = rc.alpha_meshes()
    for lv in (2, 5):
        c5 = rc.Context(rob, obs.subdivide(lv), device=0)
        med, best, nv = timeit(c5, rc.createAlphaBotConfigurations(100_000, True, 508), 5)
        rows.append((f"c5h_l{lv}_hard100k", 100_000, med, best, nv))
        mix = np.concatenate([rc.createAlphaBotConfigurations(1_125_000, False, 511),
                              rc.createAlphaBotConfigurations(125_000, True, 512)])
        med, best, nv = timeit(c5, mix, 5)
        rows.append((f"c5_l{lv}_mix1.25M", len(mix), med, best, nv))
        c5.close()
for name, n, med, best, nv in rows:
    print(json.dumps({"tag": tag, "workload": name, "ms_median": round(med, 4), "ms_best": round(best, 4),
                      "Mconfigs_per_s": round(n / med / 1e3, 1), "valid": nv}))
