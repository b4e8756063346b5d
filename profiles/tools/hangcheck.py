"""Quick liveness + agreement check of RC_MODE_CULLED against RC_MODE_BRUTE through rc_check_host
(numpy only, no torch import): used under `timeout` while developing the hand-over protocol.
usage: hangcheck.py [n_hard] [n_easy]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc

nh = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
ne = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
t0 = time.time()
ctx = rc.Context(device=0)
confs = np.concatenate([rc.createAlphaBotConfigurations(nh, True, 7),
                        rc.createAlphaBotConfigurations(ne, False, 8)]).astype(np.float32)
ctx.stats_enable(True)
for rep in range(3):
    got = ctx.check_host(confs, rc.RC_MODE_CULLED)
st = ctx.stats(reset=True)
ctx.stats_enable(False)
got2 = ctx.check_host(confs, rc.RC_MODE_CULLED)
want = ctx.check_host(confs, rc.RC_MODE_BRUTE)
print("env", {k: os.environ.get(k) for k in ("RC_SHARE", "RC_PDL")}, "n", len(confs),
      "mismatch(stats kernel)", int((got != want).sum()), "mismatch(plain kernel)", int((got2 != want).sum()),
      "shared", st["shared_pieces"], "watchdog", st["watchdog"], st["watchdog_where"],
      "kernel ms", round(ctx.last_kernel_ms(), 4), "wall %.1fs" % (time.time() - t0), flush=True)
