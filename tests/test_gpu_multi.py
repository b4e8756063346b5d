"""N>1 path on real GPUs: two ranks (torchrun-style env, NCCL), configurations sharded, the kernel
writes the bitmask words into its slot of the gather buffer, one all-gather.  Skipped with fewer
than two GPUs (the CPU/gloo twin is tests/test_dist_gloo.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import robocheck_b200 as rc
    from robocheck_b200 import sharding as sh
    ctx = rc.Context(device=rank)
    confs = np.concatenate([rc.createAlphaBotConfigurations(n // 2, True, 41),
                            rc.createAlphaBotConfigurations(n - n // 2, False, 42)])

    def check_local(shard, mask_words):
        if len(shard):
            d = torch.from_numpy(np.ascontiguousarray(shard)).cuda()
            ctx.check(d, len(shard), None, mask_words, rc.RC_MODE_CULLED, torch.cuda.current_stream())

    flags = sh.check_sharded(check_local, confs, dist, device=torch.device("cuda", rank))
    single = ctx.check_host(confs)
    ok = bool(np.array_equal(flags, single)) and len(flags) == n
    # the same gather fused behind the check, over peer-mapped memory (rc_check_gather): three calls in
    # a row (both gather areas, call numbers), ragged sizes incl. a short last shard
    fused = "skipped"
    try:
        pg = sh.PeerGather(rc, dist, torch.device("cuda", rank), sh.words_per_shard(n, world))
    except Exception as exc:  # symmetric memory not available on this box: NCCL path only
        pg, fused = None, f"unavailable: {type(exc).__name__}: {exc}"
    if pg is not None:
        fused = "ok"
        for call, m in enumerate((n, n - 4097, 70)):
            sub = confs[:m]
            lo, hi, per = sh.shard_bounds(m, world, rank)
            # a shard never exceeds the slot the buffer was sized for
            lo, hi = min(m, rank * pg.words * 32), min(m, (rank + 1) * pg.words * 32)
            d = torch.from_numpy(np.ascontiguousarray(sub[lo:hi])).cuda()
            words = pg.check(ctx, d, hi - lo, None, torch.cuda.current_stream())
            torch.cuda.synchronize()
            got = sh.unpack_bitmask(words.cpu().numpy(), world * pg.words * 32)[:m]
            if not np.array_equal(got, single[:m]):
                ok, fused = False, f"call {call}: {int((got != single[:m]).sum())} flags differ"
        if ctx.stats()["watchdog"] != 0:
            ok, fused = False, "watchdog"
    agree = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    if rank == 0:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        if os.path.isdir(out):
            with open(os.path.join(out, "fused_gather_status.txt"), "w") as f:
                f.write(fused + "\n")
        q.put(bool(agree.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_check():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 50001, q)) for r in range(2)]
    [p.start() for p in procs]
    ok = q.get(timeout=300)
    [p.join(timeout=60) for p in procs]
    assert ok and all(p.exitcode == 0 for p in procs)


def test_check_multi_one_context_per_gpu():
    """rc_check_multi with one context per device (single process): host bitmask and flags, and the
    full bitmask on every device after the peer copies, equal the single-GPU result."""
    import torch
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs two GPUs")
    import robocheck_b200 as rc
    n = 200_003
    confs = np.concatenate([rc.createAlphaBotConfigurations(n // 2, True, 51),
                            rc.createAlphaBotConfigurations(n - n // 2, False, 52)])
    confs = confs[np.random.default_rng(6).permutation(n)]
    ctxs = [rc.Context(device=d) for d in range(n_dev)]
    try:
        want = ctxs[0].check_host(confs)
        words = (n + 31) // 32
        d_all = [torch.full((words,), -1, dtype=torch.int32, device=torch.device("cuda", d)) for d in range(n_dev)]
        for d in range(n_dev):
            torch.cuda.synchronize(d)
        mask, flags = rc.check_multi(ctxs, confs, want_flags=True, d_bitmask_all=d_all)
        assert np.array_equal(flags, want)
        assert np.array_equal(np.unpackbits(mask.view(np.uint8), bitorder="little")[:n], want)
        for t in d_all:
            assert np.array_equal(t.cpu().numpy().view(np.uint32), mask)
    finally:
        for c in ctxs:
            c.close()
