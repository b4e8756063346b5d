"""Short driver for the ncu passes: a few launches of each kernel of the hot path."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robocheck_b200 as rc

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = 100_000
ctx = rc.Context(device=0)
confs = rc.createAlphaBotConfigurations(n, True, 508)
d_confs = torch.from_numpy(confs).cuda()
d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
d_mask = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream()
reps = 3
if which in ("all", "culled"):
    for _ in range(reps):
        ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, s)
if which in ("all", "brute"):
    for _ in range(reps):
        ctx.check(d_confs, 16384, d_valid, None, rc.RC_MODE_BRUTE, s)
if which in ("all", "stage"):
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
torch.cuda.synchronize()
print("ok", int(d_valid.sum()))
