"""Per-warp timeline of the global-obstacle descent kernel (BASELINE config 5, one of eight shards), from a
-DRC_TIMELINE build.  usage: RC_LIB=ab/librobocheck_tl.so python profiles/tools/timeline_c5.py [knob=value ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc  # noqa: E402

rob, obs = rc.alpha_meshes()
ctx = rc.Context(rob, obs.subdivide(5), device=0)
s = torch.cuda.current_stream()
W = 148 * 32
n = 1_250_000
hard = rc.createAlphaBotConfigurations(n // 10, True, 511)
easy = rc.createAlphaBotConfigurations(n - n // 10, False, 1511)
confs = np.concatenate([easy, hard])[np.random.default_rng(511).permutation(n)]
for name, cf in (("mix 1.25M", confs), ("hard 125k", hard)):
    m_ = len(cf)
    d = torch.from_numpy(np.ascontiguousarray(cf)).cuda()
    v = torch.zeros(m_, dtype=torch.uint8, device="cuda")
    m = torch.zeros((m_ + 31) // 32, dtype=torch.int32, device="cuda")
    tr = torch.zeros(8 * max(m_, W), dtype=torch.int32, device="cuda")
    for _ in range(2):
        ctx.check(d, m_, v, m, rc.RC_MODE_CULLED, s)
    torch.cuda.synchronize()
    ctx.stats_trace(tr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); ctx.check(d, m_, v, m, rc.RC_MODE_CULLED, s); e1.record(s)
    torch.cuda.synchronize()
    ctx.stats_trace(None)
    t = tr.cpu().numpy().view(np.uint32)[: 8 * W].reshape(W, 8).astype(np.int64)
    t = t[t[:, 0] != 0]  # 24 warps per CTA leave the other rows empty
    t0 = t[:, 0].min()
    dry, leave, ent = (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, t[:, 3]
    q = lambda x: " ".join(f"{np.percentile(x, p):8.1f}" for p in (0, 10, 50, 90, 99, 100))
    print(f"{name}: event-timed call {e0.elapsed_time(e1) * 1e3:.1f} us; percentiles 0 10 50 90 99 100 (us after the first warp started)")
    print(f"  list dry     {q(dry)}")
    print(f"  warp leaves  {q(leave)}")
    print(f"  entries/warp {q(ent)}  total {ent.sum()}  donations {int(t[:, 6].sum())}")
    order = np.argsort(-t[:, 4])[:10]
    print("  longest entries (us, node/leaf/tail rounds capped at 1023, warp leaves at): " +
          "  ".join(f"{t[w, 4] / 1e3:.0f}:{t[w, 5] >> 20}/{(t[w, 5] >> 10) & 1023}/{t[w, 5] & 1023}@{leave[w]:.0f}" for w in order))
    busy = [(x, int((leave > x).sum())) for x in np.linspace(0, leave.max(), 11)]
    print("  warps not yet gone at t: " + "  ".join(f"{x:.0f}:{c}" for x, c in busy))
