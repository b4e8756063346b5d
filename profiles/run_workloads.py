#!/usr/bin/env python
"""Every BASELINE.json configuration on one GPU (device-resident configs/s, CUDA events, best of 5):
  C1 easy_confs10,000 (shipped set)            C2 hard 100k (seed 508) -- the bench.py headline
  C4 1M "uniform SE(3)" = the reference's easy sampler (seed 510)
  C5 alpha1.0 robot vs the midpoint-subdivided obstacle: 2 levels = 16,128 triangles ("16x"),
     5 levels = 1,032,192 triangles ("~1M"); 90% easy + 10% hard configurations (seed 511)
  C3 sphere-approximated robot (this repo's definition: one enclosing sphere per robot triangle)
     vs the obstacle mesh, 1M easy configurations (seed 509) and 100k hard ones"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robocheck_b200 as rc

rob, obs = rc.alpha_meshes()
gold = np.load(os.path.join(rc.REPO_ROOT, "tests", "golden", "easy_confs10k.npz"))


def mix(n, seed):
    hard = rc.createAlphaBotConfigurations(n // 10, True, seed)
    easy = rc.createAlphaBotConfigurations(n - n // 10, False, seed + 1000)
    out = np.concatenate([easy, hard])
    return out[np.random.default_rng(seed).permutation(n)]


def run(name, ctx, confs, reps=5, spheres=False):
    n = len(confs)
    d = torch.from_numpy(np.ascontiguousarray(confs, dtype=np.float32)).cuda()
    v = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream()
    ctx.stats_enable(True); ctx.stats(reset=True)
    check = (lambda: ctx.check_spheres(d, n, v, None, s)) if spheres else \
            (lambda: ctx.check(d, n, v, None, rc.RC_MODE_CULLED, s))
    check(); torch.cuda.synchronize()
    st = ctx.stats(reset=True); ctx.stats_enable(False)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        check()
        e1.record(s); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    rec = {"workload": name, "configs": n, "ms": round(best, 4), "configs_per_s": round(n / best * 1e3),
           "valid_fraction": round(float(v.float().mean()), 4),
           "mix": {k: st[k] for k in ("sphere_cleared", "broad_cleared", "narrow_valid", "narrow_invalid", "overflow")},
           "per_config": {k: round(st[k] / n, 2) for k in ("node_pairs", "leaf_pairs", "tri_pairs", "tail_pairs")}}
    print(json.dumps(rec), flush=True)
    return v.cpu().numpy()


ctx = rc.Context(rob, obs, device=0)
flags = run("C1 easy_confs10,000.conf", ctx, gold["confs"])
assert np.array_equal(flags, gold["tags"]), "C1 must reproduce the FCL tags"
run("C2 hard 100,000 (seed 508)", ctx, rc.createAlphaBotConfigurations(100_000, True, 508))
run("C2s hard 10,000 (seed 508; a planner-sized batch)", ctx, rc.createAlphaBotConfigurations(10_000, True, 508))
run("C2t hard 2,000 (seed 508)", ctx, rc.createAlphaBotConfigurations(2_000, True, 508))
run("C4 uniform 1,000,000 (easy sampler, seed 510)", ctx, rc.createAlphaBotConfigurations(1_000_000, False, 510))
run("C2x hard 1,000,000 (seed 512)", ctx, rc.createAlphaBotConfigurations(1_000_000, True, 512))
run("C3 sphere robot (1008 spheres), easy 1,000,000 (seed 509)", ctx,
    rc.createAlphaBotConfigurations(1_000_000, False, 509), spheres=True)
run("C3h sphere robot, hard 100,000 (seed 508)", ctx, rc.createAlphaBotConfigurations(100_000, True, 508), spheres=True)
ctx.close()
for levels in (2, 5):
    t0 = time.time()
    sub = obs.subdivide(levels)
    c5 = rc.Context(rob, sub, device=0)
    build_s = time.time() - t0
    n = 1_250_000  # 10M configurations over 8 GPUs
    run(f"C5 obstacle subdivided x{levels} ({sub.n_t} triangles), {n} configs, context build {build_s:.1f} s",
        c5, mix(n, 511))
    run(f"C5h same obstacle, hard 100,000", c5, rc.createAlphaBotConfigurations(100_000, True, 508))
    c5.close()
