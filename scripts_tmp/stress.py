import os, sys, subprocess, json
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import robocheck_b200 as rc
levels, n, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rob, obs0 = rc.alpha_meshes(); obs = obs0.subdivide(levels) if levels else obs0
confs = rc.createAlphaBotConfigurations(n, True, 700 + levels)
c = rc.Context(rob, obs, device=0)
d = torch.from_numpy(confs).cuda(); s = torch.cuda.current_stream()
v = torch.zeros(n, dtype=torch.uint8, device="cuda")
c.stats_enable(True); c.check(d, n, v, None, rc.RC_MODE_CULLED, s); torch.cuda.synchronize(); c.stats_enable(False)
want = v.cpu().numpy().copy()
bad = []
for r in range(reps):
    v.fill_(7)
    c.check(d, n, v, None, rc.RC_MODE_CULLED, s); torch.cuda.synchronize()
    k = int((v.cpu().numpy() != want).sum())
    if k: bad.append((r, k))
print(os.environ.get("TAG"), "levels", levels, "n", n, "reps", reps, "bad runs", len(bad), bad[:6], "watchdog", c.stats()["watchdog"])
'''
open("/tmp/stress_child.py", "w").write(code)
for env in ({}, {"RC_SPLIT_MAX": "1"}, {"RC_SHARE": "0"}, {"RC_SPLIT_MAX": "1", "RC_SHARE": "0"}):
    for lv, n, reps in ((2, 160, 300), (0, 160, 300), (2, 3000, 100), (0, 3000, 100)):
        e = dict(os.environ); e.update(env); e["TAG"] = json.dumps(env)
        subprocess.run([sys.executable, "/tmp/stress_child.py", str(lv), str(n), str(reps)], env=e)
