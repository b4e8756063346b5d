import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc
ctx = rc.Context(device=0)
n = 100000
confs = rc.createAlphaBotConfigurations(n, True, 508)
d = torch.from_numpy(confs).cuda(); v = torch.zeros(n, dtype=torch.uint8, device='cuda')
s = torch.cuda.current_stream()
dx=dy=dz=None
vb = torch.zeros(n, dtype=torch.uint8, device='cuda')
ctx.broad_fused(d, n, None, None, None, None, vb, s)
ctx.check(d, n, v, None, rc.RC_MODE_CULLED, s); torch.cuda.synchronize()
vb = vb.cpu().numpy().astype(bool); vv = v.cpu().numpy().astype(bool)
classes = {'cleared': vb, 'narrow_valid': (~vb) & vv, 'colliding': ~vv, 'all': np.ones(n, bool)}
def timeit(c):
    idx = np.nonzero(classes[c])[0]
    sub = confs[np.resize(idx, n)]
    dd = torch.from_numpy(np.ascontiguousarray(sub)).cuda()
    for _ in range(3): ctx.check(dd, n, v, None, rc.RC_MODE_CULLED, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10): ctx.check(dd, n, v, None, rc.RC_MODE_CULLED, s)
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for c in classes:
    print(f"{c:14s} share {classes[c].mean():.3f}  {timeit(c)*1000:8.1f} us per 100k  ({timeit(c)*10:.2f} ns/config)")
ctx.stats_enable(True)
for c in classes:
    idx = np.nonzero(classes[c])[0]
    sub = confs[np.resize(idx, n)]
    dd = torch.from_numpy(np.ascontiguousarray(sub)).cuda()
    ctx.stats(reset=True)
    ctx.check(dd, n, v, None, rc.RC_MODE_CULLED, s); torch.cuda.synchronize()
    st = ctx.stats(reset=True)
    print(c, {k: round(st[k]/n,1) for k in ('node_pairs','leaf_pairs','tri_pairs','tail_pairs','node_rounds','leaf_rounds','tail_rounds')})
