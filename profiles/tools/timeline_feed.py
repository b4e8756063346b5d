"""Where the warps of the single-launch kernel (RC_FEED=2) spend their time outside the descents, from a
-DRC_TIMELINE build.  usage: RC_FEED=2 RC_LIB=ab/librobocheck_tl.so python profiles/tools/timeline_feed.py [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
s = torch.cuda.current_stream()
ctx = rc.Context(device=0)
W = 148 * 32
d = torch.from_numpy(rc.createAlphaBotConfigurations(n, True, 508)).cuda()
v = torch.zeros(n, dtype=torch.uint8, device="cuda")
m = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
for _ in range(3):
    ctx.check(d, n, v, m, rc.RC_MODE_CULLED, s)
torch.cuda.synchronize()
tr = torch.zeros(8 * max(n, W), dtype=torch.int32, device="cuda")
ctx.stats_trace(tr)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s); ctx.check(d, n, v, m, rc.RC_MODE_CULLED, s); e1.record(s)
torch.cuda.synchronize()
ctx.stats_trace(None)
t = tr.cpu().numpy().view(np.uint32)[: 8 * W].reshape(W, 8).astype(np.int64)
t = t[t[:, 0] != 0]  # 24 warps per CTA leave the other rows empty
t0 = t[:, 0].min()
leave, ent = (t[:, 2] - t0) / 1e3, t[:, 3]
q = lambda x: " ".join(f"{np.percentile(x, p):8.1f}" for p in (0, 10, 50, 90, 99, 100))
print(f"n={n}: event-timed call {e0.elapsed_time(e1) * 1e3:.1f} us; percentiles 0 10 50 90 99 100 over the warps")
print(f"  warp leaves (us)             {q(leave)}")
print(f"  entries per warp             {q(ent)}   total {ent.sum()}")
print(f"  us inside feed_next          {q(t[:, 4] / 1e3)}   mean {t[:, 4].mean() / 1e3:.1f}")
print(f"  iterations with empty ring   {q(t[:, 5])}   total {t[:, 5].sum()}")
print(f"  us inside feed_refill        {q(t[:, 6] / 1e3)}   total refills {t[:, 7].sum()}  mean us per refill {t[:, 6].sum() / max(t[:, 7].sum(), 1) / 1e3:.2f}")
