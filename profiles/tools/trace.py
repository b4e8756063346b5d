"""Per-configuration cost trace of the descent step (rc_stats_trace): distribution of the cycles a
warp spends on one configuration, when the last ones are decided, and what a perfectly balanced
schedule would take.  usage: trace.py [n] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 508
ctx = rc.Context(device=0)
confs = rc.createAlphaBotConfigurations(n, True, seed)
d_confs = torch.from_numpy(confs).cuda()
d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
d_trace = torch.zeros(4 * n, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream()
for _ in range(3):
    ctx.check(d_confs, n, d_valid, None, rc.RC_MODE_CULLED, s)
ctx.stats_enable(True)
ctx.stats_trace(d_trace)
ctx.check(d_confs, n, d_valid, None, rc.RC_MODE_CULLED, s)
torch.cuda.synchronize()
ctx.stats_trace(None)
ctx.stats_enable(False)
t = d_trace.cpu().numpy().view(np.uint32).reshape(n, 4)
if os.environ.get("RC_TRACE_DUMP"):
    np.savez_compressed(os.environ["RC_TRACE_DUMP"], trace=t, valid=d_valid.cpu().numpy(), confs=confs)
v = d_valid.cpu().numpy()
cyc, ns = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64)
listed = cyc > 0
ns0 = ns[listed].min()
fin = (ns[listed] - ns0) / 1e3
c = cyc[listed]
print(f"listed {listed.sum()} of {n}; cycles per listed configuration: mean {c.mean():.0f} median {np.median(c):.0f} "
      f"p90 {np.percentile(c, 90):.0f} p99 {np.percentile(c, 99):.0f} p99.9 {np.percentile(c, 99.9):.0f} max {c.max()}")
for name, m in (("valid", v[listed] == 1), ("colliding", v[listed] == 0)):
    cc = c[m]
    print(f"  {name:9s} {m.sum():6d}: mean {cc.mean():.0f} p99 {np.percentile(cc, 99):.0f} max {cc.max()}")
print(f"last decision {fin.max():.1f} us after the first; 50/90/99/99.9 % decided after "
      + "/".join(f"{np.percentile(fin, q):.1f}" for q in (50, 90, 99, 99.9)) + " us")
warps = 148 * 32
print(f"sum of cycles / {warps} warps = {c.sum() / warps:.0f} cycles = {c.sum() / warps / 1.965e3:.1f} us at 1965 MHz; "
      f"longest single configuration {c.max() / 1.965e3:.1f} us")
order = np.argsort(fin)[-10:]
print("last ten decided: cycles", c[order].tolist(), "at us", np.round(fin[order], 1).tolist())
nr, lr, tr, ms = (t[listed, 2] >> 16), (t[listed, 2] & 0xffff), (t[listed, 3] >> 16), (t[listed, 3] & 0xffff)
print("last ten: node rounds", nr[order].tolist(), "leaf rounds", lr[order].tolist(), "tail rounds", tr[order].tolist(),
      "deepest stack", ms[order].tolist())
heavy = c > np.percentile(c, 99)
print(f"heaviest 1 %: node rounds {nr[heavy].mean():.0f} leaf rounds {lr[heavy].mean():.0f} tail rounds {tr[heavy].mean():.0f} "
      f"deepest stack mean {ms[heavy].mean():.0f}; all listed: {nr.mean():.1f} / {lr.mean():.1f} / {tr.mean():.1f} / {ms.mean():.0f}")
A = np.stack([nr, lr, tr, np.ones_like(nr)], 1).astype(np.float64)
coef = np.linalg.lstsq(A, c.astype(np.float64), rcond=None)[0]
print("cycles ~ %.0f x node rounds + %.0f x leaf rounds + %.0f x tail rounds + %.0f" % tuple(coef))
st = ctx.stats()
print({k: st[k] for k in ("node_pairs", "node_rounds", "leaf_pairs", "leaf_rounds", "tail_pairs", "tail_rounds")})
