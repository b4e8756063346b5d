"""Per-warp timeline of check_culled_kernel on planner-sized batches, from a -DRC_TIMELINE build of the
library: when every warp started (tables staged), found the list dry, and left.
usage: RC_LIB=ab/librobocheck_tl.so python profiles/tools/timeline.py [n ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [2000, 10000, 100000]
s = torch.cuda.current_stream()
ctx = rc.Context(device=0)
W = 148 * 32
for n in sizes:
    confs = rc.createAlphaBotConfigurations(n, True, 508)
    d = torch.from_numpy(confs).cuda()
    v = torch.zeros(n, dtype=torch.uint8, device="cuda")
    m = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    tr = torch.zeros(8 * max(n, W), dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.check(d, n, v, m, rc.RC_MODE_CULLED, s)
    torch.cuda.synchronize()
    ctx.stats_trace(tr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); ctx.check(d, n, v, m, rc.RC_MODE_CULLED, s); e1.record(s)
    torch.cuda.synchronize()
    ctx.stats_trace(None)
    t = tr.cpu().numpy().view(np.uint32)[: 8 * W].reshape(W, 8).astype(np.int64)
    t = t[t[:, 0] != 0]  # calls below 65,536 configurations launch 24 or 16 warps per CTA: the other rows stay empty
    t0 = t[:, 0].min()
    start, dry, leave, ent = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, t[:, 3]
    q = lambda x: " ".join(f"{np.percentile(x, p):7.1f}" for p in (0, 10, 50, 90, 99, 100))
    print(f"n={n}  event-timed call {e0.elapsed_time(e1) * 1e3:.1f} us; us after the first warp started, percentiles 0 10 50 90 99 100")
    print(f"  warp start   {q(start)}")
    print(f"  list dry     {q(dry)}")
    print(f"  warp leaves  {q(leave)}")
    print(f"  entries/warp {q(ent)}   total {ent.sum()}")
    busy_after = [(x, int((dry > x).sum())) for x in np.linspace(0, leave.max(), 9)]
    print("  warps still drawing from the list at t: " + "  ".join(f"{x:.0f}us:{c}" for x, c in busy_after))
    print(f"  donations made {int(t[:, 6].sum())}, pieces received {int(t[:, 7].sum())}")
    order = np.argsort(-t[:, 4])[:12]
    print("  longest entries (us, node/leaf/tail rounds, warp's leave time): " +
          "  ".join(f"{t[w, 4] / 1e3:.1f}:{t[w, 5] >> 20}/{(t[w, 5] >> 10) & 1023}/{t[w, 5] & 1023}@{leave[w]:.0f}" for w in order))
    late = np.argsort(-leave)[:12]
    print("  last warps to leave (leave us, dry us, entries, longest entry us, rounds): " +
          "  ".join(f"{leave[w]:.0f},{dry[w]:.0f},{ent[w]},{t[w, 4] / 1e3:.1f},{t[w, 5] >> 20}/{(t[w, 5] >> 10) & 1023}/{t[w, 5] & 1023}" for w in late))
