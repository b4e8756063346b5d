"""world_size-2 test of the N>1 host path on CPU (gloo): shard -> local check -> bitmask
all-gather.  The local check is the CPU oracle here (no GPU in this container); on the GPU box
the same check_sharded() drives Context.check (tests/test_gpu_multi.py, bench.py)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import robocheck_b200 as rc
    from oracle import pyoracle as po
    from robocheck_b200 import sharding as sh
    rob_m, obs_m = rc.alpha_meshes()
    rob, obs = po.OracleMesh(*rob_m.arrays()), po.OracleMesh(*obs_m.arrays())
    confs = np.concatenate([rc.createAlphaBotConfigurations(n // 2, True, 31),
                            rc.createAlphaBotConfigurations(n - n // 2, False, 32)])

    def check_local(shard, mask_words):
        flags = po.check(rob, obs, shard, po.VARIANT_GPU, nthreads=2) if len(shard) else np.zeros(0, np.uint8)
        mask_words.copy_(torch.from_numpy(sh.pack_bitmask(flags, len(mask_words)).view(np.int32)))

    flags = sh.check_sharded(check_local, confs, dist)
    if rank == 0:
        want = po.check(rob, obs, confs, po.VARIANT_GPU, nthreads=2)
        q.put((flags.tolist(), want.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [150, 64])
def test_sharded_check_two_ranks(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    got, want = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert got == want and len(got) == n
