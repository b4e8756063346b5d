import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robocheck_b200 as rc
n = int(sys.argv[1]); hard = sys.argv[2] == "hard"
ctx = rc.Context(device=0)
confs = rc.createAlphaBotConfigurations(n, hard, 508)
if len(sys.argv) > 3: confs[:, :3] += 5000.0
d = torch.from_numpy(confs).cuda()
v = torch.zeros(n, dtype=torch.uint8, device="cuda")
m = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream()
ctx.check(d, n, v, m, rc.RC_MODE_CULLED, s)
torch.cuda.synchronize()
print("ok", n, hard, int(v.sum()), ctx.stats())
