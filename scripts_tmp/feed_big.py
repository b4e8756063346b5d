import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robocheck_b200 as rc
ctx = rc.Context(device=0)
for n, hard, seed in ((1_000_000, True, 512), (3_000_000, False, 510), (1_500_000, True, 77)):
    confs = rc.createAlphaBotConfigurations(n, hard, seed)
    if not hard:
        confs[::7] = rc.createAlphaBotConfigurations((n + 6) // 7, True, seed + 1)[: len(confs[::7])]
    pinned = torch.from_numpy(confs).pin_memory()
    d = torch.from_numpy(confs).cuda()
    v = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ctx.check(d, n, v, None, rc.RC_MODE_CULLED, torch.cuda.current_stream())
    torch.cuda.synchronize()
    want = v.cpu().numpy()
    for rep in range(3):
        t0 = time.perf_counter()
        got = ctx.check_host(pinned.numpy())
        dt = time.perf_counter() - t0
        assert np.array_equal(got, want), (n, rep, int((got != want).sum()))
    print(f"n={n} hard={hard}: host call (single launch) == device path on every flag, {n / dt / 1e6:.1f} M configs/s, valid {int(want.sum())}, watchdog {ctx.stats()['watchdog']}", flush=True)
