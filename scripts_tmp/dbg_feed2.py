import os, sys, threading, time
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
tr = torch.zeros(32 * 16, dtype=torch.int32).pin_memory()
ctx.stats_trace(tr)
s = torch.cuda.current_stream()
def dump():
    time.sleep(8)
    t = tr.numpy().reshape(32, 16)
    for w in range(32):
        if t[w][0] != 7032: print("warp", w, t[w].tolist(), flush=True)
    os._exit(0)
threading.Thread(target=dump, daemon=True).start()
try:
    ctx.check(d, n, v, m, rc.RC_MODE_CULLED, s)
    torch.cuda.synchronize()
    print("ok", n, hard, int(v.sum()))
except Exception as e:
    print("ERR", str(e)[:100])
t = tr.numpy().reshape(32, 16)
for w in range(32):
    if t[w][0] != 7032: print("warp", w, t[w].tolist(), flush=True)
