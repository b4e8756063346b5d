import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc
ctx = rc.Context(device=0)
n = 65536
confs = rc.createAlphaBotConfigurations(n, True, 508)
d = torch.from_numpy(confs).cuda(); v = torch.zeros(n, dtype=torch.uint8, device='cuda')
s = torch.cuda.current_stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best=1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); fn(); e1.record(s); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
ctx.stats_enable(True); ctx.stats(reset=True)
ctx.check(d, n, v, None, rc.RC_MODE_BRUTE, s); torch.cuda.synchronize()
st = ctx.stats(reset=True); ctx.stats_enable(False)
ms = t(lambda: ctx.check(d, n, v, None, rc.RC_MODE_BRUTE, s))
pairs = st['narrow_valid']*1008*1008 + st['narrow_invalid']
peak = ctx.fp32_peak_tflops()
print(f"brute {n} configs {ms:.2f} ms  {n/ms*1e3/1e6:.3f}M configs/s  {18*pairs/ms/1e9:.2f} TFLOP/s alg  frac {18*pairs/ms/1e9/peak:.3f} of {peak:.1f}")
# write-only ceiling
buf = torch.empty(2<<30, dtype=torch.uint8, device='cuda')
ms = t(lambda: buf.fill_(3), 5)
print(f"torch fill_ 2 GiB: {buf.numel()/ms/1e6:.0f} GB/s write-only")
src = torch.empty(1<<30, dtype=torch.uint8, device='cuda'); dst = torch.empty(1<<30, dtype=torch.uint8, device='cuda')
ms = t(lambda: dst.copy_(src), 5)
print(f"torch copy 1 GiB: {2*src.numel()/ms/1e6:.0f} GB/s read+write")
del buf, src, dst
nn = 100000; nv = ctx.n_rob_v
dd = torch.from_numpy(rc.createAlphaBotConfigurations(nn, True, 508)).cuda()
dx, dy, dz = (torch.empty(nn*nv, dtype=torch.float32, device='cuda') for _ in range(3))
box = torch.empty(nn*6, dtype=torch.float32, device='cuda'); vv = torch.zeros(nn, dtype=torch.uint8, device='cuda')
ms = t(lambda: ctx.broad_fused(dd, nn, dx, dy, dz, box, vv, s), 5)
print(f"transform stage: {ms*1e3:.1f} us  {9553*nn/ms/1e6:.0f} GB/s algorithmic")
ms = t(lambda: ctx.broad_fused(dd, nn, None, None, None, box, vv, s), 5)
print(f"transform stage without vertex store: {ms*1e3:.1f} us")
aos = torch.empty(nn*nv*3, dtype=torch.float32, device='cuda')
ms = t(lambda: ctx.transform_aos(dd, nn, aos, stream=s), 5)
print(f"transform AoS stage: {ms*1e3:.1f} us  {(24+9504)*nn/ms/1e6:.0f} GB/s")
ms = t(lambda: ctx.generate_aabb(aos, nv, nn, box, s), 5)
print(f"AABB AoS stage: {ms*1e3:.1f} us  {(9504+24)*nn/ms/1e6:.0f} GB/s")
ms = t(lambda: ctx.generate_aabb_soa(dx, dy, dz, nv, nn, box, s), 5)
print(f"AABB SoA stage: {ms*1e3:.1f} us  {(9504+24)*nn/ms/1e6:.0f} GB/s")
ms = t(lambda: ctx.narrow_phase(dx, dy, dz, 16384, vv, s), 2)
print(f"narrow stage (16384 configs, exhaustive): {ms:.2f} ms")
