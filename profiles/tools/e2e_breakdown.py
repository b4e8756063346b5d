"""Where the synchronous host call spends its time on this box: wall time of rc_check_host against its
parts (H2D copy + sync, the two kernels device-resident + sync, CUDA-event time of the kernels inside the
call) and against the pipelined calls.  100,000 hard configurations, medians of 60."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import robocheck_b200 as rc

n = 100_000
ctx = rc.Context(device=0)
confs = rc.createAlphaBotConfigurations(n, True, 508)
h_confs = torch.from_numpy(confs).pin_memory()
h_valid = torch.zeros(n, dtype=torch.uint8).pin_memory()
h_valid2 = torch.zeros(n, dtype=torch.uint8).pin_memory()
npc, npv, npv2 = h_confs.numpy(), h_valid.numpy(), h_valid2.numpy()
d_confs = torch.empty_like(h_confs, device="cuda")
d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
lib = rc.lib()


def med(f, reps=60):
    for _ in range(5):
        f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e6, ts[0] * 1e6


def full():
    assert lib.rc_check_host(ctx._h, npc.ctypes.data, n, npv.ctypes.data, rc.RC_MODE_CULLED) == 0


def h2d():
    d_confs.copy_(h_confs, non_blocking=True)
    torch.cuda.synchronize()


def kernels():
    ctx.check(d_confs, n, d_valid, None, rc.RC_MODE_CULLED, s)
    torch.cuda.synchronize()


def sync_only():
    torch.cuda.synchronize()


kms = []


def full_with_kernel_time():
    full()
    kms.append(float(lib.rc_last_kernel_ms(ctx._h)) * 1e3)


lib.rc_last_kernel_ms.restype = __import__("ctypes").c_float
print("rc_check_host (synchronous)      median %.1f us  min %.1f us" % med(full))
med(full_with_kernel_time)
kms.sort()
print("  kernels inside the call (CUDA events) median %.1f us" % kms[len(kms) // 2])
print("H2D 2.4 MB + sync                median %.1f us  min %.1f us" % med(h2d))
print("kernels device-resident + sync   median %.1f us  min %.1f us" % med(kernels))
print("empty synchronize                median %.1f us  min %.1f us" % med(sync_only))
for mode, env in (("direct flag stores", "1"), ("device-to-host copy", "0")):
    os.environ["RC_HOST_DIRECT"] = env
    c2 = rc.Context(device=0)

    def f2():
        assert lib.rc_check_host(c2._h, npc.ctypes.data, n, npv.ctypes.data, rc.RC_MODE_CULLED) == 0
    print("rc_check_host, %-20s median %.1f us  min %.1f us" % ((mode,) + med(f2)))
    c2.close()
K = 40


def pipelined():
    prev = ctx.check_host_submit(npc, npv)
    for k in range(1, K):
        cur = ctx.check_host_submit(npc, npv2 if k & 1 else npv)
        ctx.check_host_wait(prev)
        prev = cur
    ctx.check_host_wait(prev)


m, lo = med(pipelined, 8)
print("pipelined, per batch             median %.1f us  min %.1f us" % (m / K, lo / K))
print("rc_check_host (first context, again at the end) median %.1f us  min %.1f us" % med(full))
med(full_with_kernel_time)
print("  kernels inside the call, last 60: median %.1f us" % sorted(kms[-60:])[30])
