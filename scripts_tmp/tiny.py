import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import robocheck_b200 as rc
ctx = rc.Context(device=0)
s = torch.cuda.current_stream()
for name, n, hard, seed in (("easy2k", 2000, False, 510), ("hard2k", 2000, True, 508), ("hard10k", 10000, True, 508)):
    confs = rc.createAlphaBotConfigurations(n, hard, seed)
    d = torch.from_numpy(confs).cuda(); v = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for _ in range(4):
        ctx.check(d, n, v, None, rc.RC_MODE_CULLED, s)
    torch.cuda.synchronize()
print("ok")
