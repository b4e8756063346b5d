"""Short driver for the ncu passes: a few launches of each kernel of the hot path.
usage: prof_driver.py [all|culled|brute|stage|gobs|spheres|planner|gather|short] ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robocheck_b200 as rc

which = set(sys.argv[1:]) or {"all"}


def on(name):
    return "all" in which or name in which


n = 100_000
ctx = rc.Context(device=0)
confs = rc.createAlphaBotConfigurations(n, True, 508)
d_confs = torch.from_numpy(confs).cuda()
d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
d_mask = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream()
reps = 2
if on("culled"):
    for _ in range(reps):
        ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, s)
if "short" in which:
    # planner-sized calls: 2,000 configurations (16 warps per CTA, every configuration cut into parts) and
    # 30,000 (24 warps per CTA) -- bound by their longest descents, not by throughput
    for m in (2_000, 30_000):
        for _ in range(reps):
            ctx.check(d_confs, m, d_valid, d_mask, rc.RC_MODE_CULLED, s)
if on("brute"):
    for _ in range(reps):
        ctx.check(d_confs, 16384, d_valid, None, rc.RC_MODE_BRUTE, s)
if on("stage"):
    nv = ctx.n_rob_v
    dx, dy, dz = (torch.empty(n * nv, dtype=torch.float32, device="cuda") for _ in range(3))
    d_box = torch.empty(n * 6, dtype=torch.float32, device="cuda")
    for _ in range(reps):
        ctx.broad_fused(d_confs, n, dx, dy, dz, d_box, d_valid, s)
    aos = torch.empty(n * nv * 3, dtype=torch.float32, device="cuda")
    ctx.transform_aos(d_confs, n, aos, stream=s)
    ctx.generate_aabb(aos, nv, n, d_box, s)
    ctx.generate_aabb_soa(dx, dy, dz, nv, n, d_box, s)
    ctx.broad_phase(n, d_box, d_valid, stream=s)
    ctx.narrow_phase(dx, dy, dz, 8192, d_valid, s)
    del dx, dy, dz, aos
if on("gobs"):
    # BASELINE config 5: obstacle subdivided five times (1,032,192 triangles, tables in global memory)
    rob, obs = rc.alpha_meshes()
    c5 = rc.Context(rob, obs.subdivide(5), device=0)
    m = 20_000
    for _ in range(reps):
        c5.check(d_confs, m, d_valid, d_mask, rc.RC_MODE_CULLED, s)
    torch.cuda.synchronize()
    c5.close()
if on("spheres"):
    for _ in range(reps):
        ctx.check_spheres(d_confs, n, d_valid, d_mask, s)
if on("planner"):
    # on-device generation, check, compaction of the valid ones
    d_gen = torch.empty(n * 6, dtype=torch.float32, device="cuda")
    d_idx = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for _ in range(reps):
        ctx.generate_device_mix(d_gen, 0, n, rc.alpha_box(False), rc.alpha_box(True), 1024, 10, 511, s)
        ctx.check(d_gen, n, None, d_mask, rc.RC_MODE_CULLED, s)
        ctx.compact(n, d_idx, d_cnt, d_bitmask=d_mask, stream=s)
if on("gather"):
    # the all-gather of the bitmask by the last CTA of the check kernel: two ranks emulated on one
    # device with the two-phase calls (kernels that wait for each other must not share a GPU)
    from robocheck_b200 import sharding as sh
    world, words = 2, sh.words_per_shard(n, 2)
    total = int(rc.lib().rc_gather_buffer_words(world, words))
    bufs = [torch.zeros(total, dtype=torch.int32, device="cuda") for _ in range(world)]
    ptrs = [b.data_ptr() for b in bufs]
    ranks = [ctx, rc.Context(device=0)]
    half = words * 32
    for _ in range(reps):
        for r in range(world):
            lo, hi = min(n, r * half), min(n, (r + 1) * half)
            ranks[r].check_gather_begin(d_confs[lo:], hi - lo, None, ptrs, r, words, s)
        for r in range(world):
            ranks[r].gather_end(ptrs, r, words, s)
    torch.cuda.synchronize()
    ranks[1].close()
torch.cuda.synchronize()
print("ok", int(d_valid.sum()))
