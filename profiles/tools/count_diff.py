"""Hit-pair-count audit (rc_hit_counts): where the tree descent and the exhaustive sweep count different
numbers of reported pairs, list the configurations and, from the CPU oracle, every reported pair of each
with the gap between the two triangles' boxes (a reported pair with boxes apart is a rounding artefact of
the reference arithmetic on nearly coplanar planes).  usage: count_diff.py [n] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 508
ctx = rc.Context(device=0)
rob_m, obs_m = rc.alpha_meshes()
R, O = po.OracleMesh(*rob_m.arrays()), po.OracleMesh(*obs_m.arrays())
confs = rc.createAlphaBotConfigurations(n, True, seed)
d = torch.from_numpy(confs).cuda()
s = torch.cuda.current_stream()
cb = torch.zeros(n, dtype=torch.int32, device="cuda")
cc = torch.zeros(n, dtype=torch.int32, device="cuda")
fb = torch.zeros(n, dtype=torch.uint8, device="cuda")
fc = torch.zeros(n, dtype=torch.uint8, device="cuda")
ctx.hit_counts(d, n, cb, rc.RC_MODE_BRUTE, fb, s)
ctx.hit_counts(d, n, cc, rc.RC_MODE_CULLED, fc, s)
torch.cuda.synchronize()
cb, cc = cb.cpu().numpy().view(np.uint32), cc.cpu().numpy().view(np.uint32)
fb, fc = fb.cpu().numpy(), fc.cpu().numpy()
diff = np.nonzero(cb != cc)[0]
print(f"{n} configurations (seed {seed}): hit-pair counts differ on {len(diff)}, flags on {int((fb != fc).sum())}; "
      f"reported pairs in total: exhaustive {int(cb.sum())}, tree {int(cc[cc != 0xffffffff].sum())}")
ov = np.stack([O.x, O.y, O.z], 1)
o9 = np.concatenate([ov[O.v1], ov[O.v2], ov[O.v3]], 1).astype(np.float32)
olo, ohi = o9.reshape(-1, 3, 3).min(1), o9.reshape(-1, 3, 3).max(1)
for i in diff[:12]:
    _, oc = po.check_counts(R, O, confs[i:i + 1])
    wx, wy, wz = po.transform(confs[i:i + 1], R.x, R.y, R.z)
    w = np.stack([wx[0], wy[0], wz[0]], 1)
    r9 = np.concatenate([w[R.v1], w[R.v2], w[R.v3]], 1).astype(np.float32)
    gaps = []
    for j in range(len(r9)):
        rlo, rhi = r9[j].reshape(3, 3).min(0), r9[j].reshape(3, 3).max(0)
        # only pairs that can pass the first plane test matter; prefilter cheaply by the robot plane
        for k in range(len(o9)):
            if po.pair_hits(r9[j], o9[k])[0]:
                gaps.append(float(np.maximum(rlo - ohi[k], olo[k] - rhi).max()))
    gaps = np.array(gaps)
    print(f"config {i}: exhaustive {cb[i]} tree {cc[i]} oracle {int(oc[0])} flag {fb[i]}/{fc[i]}; reported pairs with "
          f"boxes apart: {int((gaps > 0).sum())} (largest gap {gaps.max():.4f}), apart by more than 2e-3: {int((gaps > 2e-3).sum())}")
