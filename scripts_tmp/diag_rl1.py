import os, sys, subprocess, json
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import robocheck_b200 as rc
from oracle import pyoracle as po
levels, n = int(sys.argv[1]), int(sys.argv[2])
rob, obs0 = rc.alpha_meshes(); obs = obs0.subdivide(levels)
confs = rc.createAlphaBotConfigurations(n, True, 700 + levels)
want = po.check(po.OracleMesh(*rob.arrays()), po.OracleMesh(*obs.arrays()), confs, po.VARIANT_GPU)
c = rc.Context(rob, obs, device=0)
d = torch.from_numpy(confs).cuda(); v = torch.zeros(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
res = {}
c.check(d, n, v, None, rc.RC_MODE_CULLED, s); torch.cuda.synchronize()
res["prod"] = int((v.cpu().numpy() != want).sum())
c.stats_enable(True); c.check(d, n, v, None, rc.RC_MODE_CULLED, s); torch.cuda.synchronize(); c.stats_enable(False)
res["stats"] = int((v.cpu().numpy() != want).sum())
st = c.stats(reset=True); res["overflow"] = st["overflow"]
bad = np.nonzero(v.cpu().numpy() != want)[0][:5].tolist()
print(os.environ.get("TAG"), levels, n, res, bad)
'''
open("/tmp/diag_rl1_child.py", "w").write(code)
for env in ({}, {"RC_SPLIT_MAX": "1"}, {"RC_SHARE": "0"}, {"RC_GMAIL": "-1"}, {"RC_CULLED_WARPS": "16"}, {"RC_CULLED_WARPS": "24"}):
    for lv, n in ((1, 400), (2, 160), (2, 2000)):
        e = dict(os.environ); e.update(env); e["TAG"] = json.dumps(env)
        subprocess.run([sys.executable, "/tmp/diag_rl1_child.py", str(lv), str(n)], env=e)
