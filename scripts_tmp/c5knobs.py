import os, sys, subprocess, json
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import robocheck_b200 as rc
rob, obs0 = rc.alpha_meshes(); obs = obs0.subdivide(5)
c = rc.Context(rob, obs, device=0)
s = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for n in (1_250_000, 10_000_000):
    d = torch.empty(n * 6, dtype=torch.float32, device="cuda")
    c.generate_device_mix(d, 0, n, rc.alpha_box(False), rc.alpha_box(True), 1024, 10, 511, s)
    m = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    for _ in range(2): c.check(d, n, None, m, rc.RC_MODE_CULLED, s)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); c.check(d, n, None, m, rc.RC_MODE_CULLED, s); e1.record(s); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out.append((n, round(min(ts), 3)))
    del d, m
print(os.environ.get("TAG"), out, "8x1.25M vs 10M speed-up bound: %.2f" % (out[1][1] / out[0][1]))
'''
open("/tmp/c5k.py", "w").write(code)
for env in ({}, {"RC_CULLED_WARPS": "32"}, {"RC_CULLED_WARPS": "24"}, {"RC_CULLED_WARPS": "16"},
            {"RC_DONATE_MIN": "64"}, {"RC_DONATE_MIN": "128"}, {"RC_MAIL_SLOTS": "8"}, {"RC_MAIL_SLOTS": "16"},
            {"RC_MAIL_SLOTS": "16", "RC_DONATE_MIN": "128"}, {"RC_TAIL_THRESH": "32"}, {"RC_TAIL_THRESH": "8"},
            {"RC_SHARE": "0"}, {}):
    e = dict(os.environ); e.update(env); e["TAG"] = json.dumps(env)
    subprocess.run([sys.executable, "/tmp/c5k.py"], env=e, timeout=200)
