"""Latency of the synchronous host call on planner-sized batches (pinned buffers, wall clock, median of 200).
usage: [RC_FEED=0|1] python profiles/tools/host_small.py [tag]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "RC_FEED=" + os.environ.get("RC_FEED", "1")
ctx = rc.Context(device=0)
lib = rc.lib()
out = {"tag": tag}
for n in (64, 512, 2000, 10000, 30000, 100000):
    confs = torch.from_numpy(rc.createAlphaBotConfigurations(n, True, 508)).pin_memory()
    valid = torch.zeros(n, dtype=torch.uint8).pin_memory()
    a, v = confs.numpy(), valid.numpy()
    for _ in range(5):
        assert lib.rc_check_host(ctx._h, a.ctypes.data, n, v.ctypes.data, rc.RC_MODE_CULLED) == 0
    ts = []
    for _ in range(200 if n <= 10000 else 50):
        t0 = time.perf_counter()
        lib.rc_check_host(ctx._h, a.ctypes.data, n, v.ctypes.data, rc.RC_MODE_CULLED)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    out[f"n{n}_us"] = round(ts[len(ts) // 2] * 1e6, 1)
    out[f"n{n}_valid"] = int(v.sum())
print(json.dumps(out), flush=True)
