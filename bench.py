#!/usr/bin/env python
"""bench.py -- collision-checked configurations per second on the alpha1.0 puzzle.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference CPU path)

Workload (config.workload): BASELINE.json configs[1], "alpha1.0 hard_confs100,000": 100,000
near-contact configurations drawn like the reference's generator (Utils_rai.cu:111-136, 322-348;
the shipped 100k file is a missing blob, so the set is regenerated with seed 508 + rank).  A
"step" is one pass of the hot path over the batch.  With N GPUs every rank checks its own 100,000
configurations (weak scaling) and the per-configuration validity bitmask is gathered on every rank
inside the timed region (by the check kernel itself over peer-mapped memory, or NCCL);
value = all configurations of all ranks / max-over-ranks device time.

Behind the headline the same run times the two north_star scaling workloads, ONE global set
sharded over the N ranks (strong scaling): scale_c4 = BASELINE configs[3] (1 M configurations) and
scale_c5 = configs[4] (10 M configurations against the 1,032,192-triangle obstacle, generated on
the device per shard), each with the gather inside the timed region.

Timing: CUDA events on the launching stream around every step, L2 flushed (256 MiB memset) before
each timed step outside the event pair, barrier + synchronize on both sides of the K steps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CONFS = 100_000
PAIRS_PER_CONFIG = 1008 * 1008
WORKLOAD = "alpha1.0 hard_confs100,000 (792 v / 1008 t robot vs 792 v / 1008 t obstacle, seed 508)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--confs", type=int, default=N_CONFS, help=argparse.SUPPRESS)
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1: bitmask all-gather fused behind the check over peer-mapped memory "
                         "(rc_check_gather), or the NCCL all-gather; auto = peer when available")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip cpu_baseline / brute / stage rooflines (used by the ncu passes)")
    ap.add_argument("--no-scale", action="store_true",
                    help="skip the sharded C4 / C5 workloads (scale_c4, scale_c5)")
    ap.add_argument("--preheat", type=float, default=1.2,
                    help="seconds of untimed steps right before the timed ones (clock sampling)")
    ap.add_argument("--c4", type=int, default=1_000_000, help=argparse.SUPPRESS)
    ap.add_argument("--c5", type=int, default=10_000_000, help=argparse.SUPPRESS)
    ap.add_argument("--c5-levels", type=int, default=5, help=argparse.SUPPRESS)
    ap.add_argument("--scale-steps", type=int, default=3, help=argparse.SUPPRESS)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                              ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        # the busy samples are the ones above the idle clock
        busy = [v for v in sm if v > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_throughput(rc, po, confs, seconds=12.0, threads=None):
    """The reference's CPU statement of the path on the host cores: rigid transform + AABB +
    broad-phase gate (oracle restatement; the reference has no CPU transform outside FCL), then
    the reference's own narrowPhaseBaseline compiled from its sources (oracle/_ref, kind
    "reference") or, when that library is absent, the oracle port of it (kind "port")."""
    import numpy as np
    threads = threads or os.cpu_count() or 1
    rob_m, obs_m = rc.alpha_meshes()
    rob, obs = po.OracleMesh(*rob_m.arrays()), po.OracleMesh(*obs_m.arrays())
    obs_box = po.aabb(obs.x, obs.y, obs.z)
    kind = "reference" if po.ref_available() else "port"

    def run(sample):
        t0 = time.perf_counter()
        if kind == "reference":
            ox, oy, oz = po.transform(sample, rob.x, rob.y, rob.z)
            lo = np.stack([ox.min(1), oy.min(1), oz.min(1)], 1)
            hi = np.stack([ox.max(1), oy.max(1), oz.max(1)], 1)
            hit = np.all((obs_box[:3] <= hi) & (lo <= obs_box[3:]), axis=1)
            flags = np.ones(len(sample), np.uint8)
            idx = np.nonzero(hit)[0]
            if len(idx):
                flags[idx] = po.ref_cpu_narrow(rob, obs, ox[idx], oy[idx], oz[idx], threads)
        else:
            flags = po.check(rob, obs, sample, po.VARIANT_HOST, threads)
        return time.perf_counter() - t0, flags

    probe = min(len(confs), 32 * threads)
    dt, _ = run(confs[:probe])
    n = int(min(len(confs), max(probe, probe * seconds / max(dt, 1e-6))))
    dt, flags = run(confs[:n])
    return {"value": n / dt, "unit": "configs/s", "cores": threads, "kind": kind,
            "sample": f"first {n} of the {len(confs)} workload configurations, {dt:.1f} s; "
                      "oracle transform+AABB+broad gate, then narrowPhaseBaseline on the rest",
            "valid_fraction": float(flags.mean())}, n, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import robocheck_b200 as rc
    from oracle import pyoracle as po
    confs = rc.createAlphaBotConfigurations(args.confs, True, 508)
    total_steps = max(1, args.steps + args.warmup)
    budget = min(12.0, 150.0 / total_steps)
    for _ in range(args.warmup):
        cpu_reference_throughput(rc, po, confs, seconds=budget)
    vals, info = [], None
    for _ in range(args.steps):
        info, n, dt = cpu_reference_throughput(rc, po, confs, seconds=budget)
        vals.append((n, dt))
    n_tot, t_tot = sum(v[0] for v in vals), sum(v[1] for v in vals)
    value = n_tot / t_tot
    info["value"] = value
    emit({
        "impl": "reference", "metric": "collision-checked configs/sec", "value": value,
        "unit": "configs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "configs_per_gpu": args.confs,
                   "note": "CPU path, bounded sample per step: " + info["sample"]},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": "configs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else any library prints on fd 1
    (NCCL's version banner, for one) has been diverted to stderr."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def divert_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


# executed floating-point work of the descent kernel, per counted event (FMA = 2; abs / negate are
# operand modifiers; DESIGN 5.1 lists the operations behind each figure)
FLOP_BOX_TEST = 78        # boxes_overlap: six axes
FLOP_TRI_SETUP = 83       # robot triangle of a leaf round: three vertices transformed (63) + its plane (20)
FLOP_TRI_BOX = 18         # triangle-box prefilter of one pair
FLOP_PLANE_TEST = 24      # three signed distances + the straddle test
FLOP_TAIL = 220           # a pair that passed both plane tests: transform, planes, distances, pair_tail


def executed_flop(st):
    return (FLOP_BOX_TEST * st["node_pairs"] + FLOP_TRI_SETUP * 4 * st["leaf_pairs"] +
            FLOP_TRI_BOX * st["tri_pairs"] + FLOP_PLANE_TEST * (st["prefilter_pairs"] + st["plane1_pairs"]) +
            FLOP_TAIL * st["tail_pairs"])


class Sharded:
    """One workload sharded over the ranks: every rank checks its contiguous shard, the validity
    bitmask is gathered inside the check (rc_check_gather over peer-mapped memory) or, when symmetric
    memory is not available, with one NCCL all-gather behind it."""

    def __init__(self, rc, dist, ctx, dev, d_confs, n_local, words, world, rank, mode, stream, d_valid=None):
        import torch
        self.rc, self.dist, self.ctx, self.d_confs, self.n, self.words = rc, dist, ctx, d_confs, n_local, words
        self.d_valid = d_valid   # optional flag bytes of the rank's own configurations, next to the bitmask
        self.world, self.rank, self.stream = world, rank, stream
        self.pg, self.note = None, "none"
        self.d_gather = torch.zeros(world * words, dtype=torch.int32, device=dev)
        self.d_mask = self.d_gather[rank * words:(rank + 1) * words]
        if world > 1:
            self.note = "all_gather of the validity bitmask (NCCL)"
            if mode != "nccl":
                try:
                    from robocheck_b200 import sharding as sh
                    self.pg = sh.PeerGather(rc, dist, dev, words)
                except Exception as exc:
                    self.pg = None
                    print(f"[bench] rank {rank}: peer gather unavailable ({type(exc).__name__}: {exc})", file=sys.stderr)
                okf = torch.tensor([1 if self.pg is not None else 0], device=dev)
                dist.all_reduce(okf, op=dist.ReduceOp.MIN)
                if int(okf.item()) == 0:
                    assert mode != "peer", "--gather peer: symmetric memory not available"
                    self.pg = None
                else:
                    self.note = ("bitmask words stored into the peers' buffers over NVLink by the last CTA of the "
                                 "check kernel (rc_check_gather, torch symmetric memory): no communicator call, "
                                 "no extra launch")

    def step(self):
        if self.pg is not None:
            return self.pg.check(self.ctx, self.d_confs, self.n, self.d_valid, self.stream)
        if self.n > 0:
            self.ctx.check(self.d_confs, self.n, self.d_valid, self.d_mask, self.rc.RC_MODE_CULLED, self.stream)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.d_gather, self.d_mask)
        return self.d_gather

    def verify_gather(self, mode):
        """Untimed: the gathered bitmask of one step against an NCCL all-gather of the same words on every
        rank; a peer gather that disagrees (or whose wait gave up) is not used for the timed steps."""
        import torch
        if self.world == 1:
            return None
        got = self.step().clone()
        if self.n > 0:
            self.ctx.check(self.d_confs, self.n, None, self.d_mask, self.rc.RC_MODE_CULLED, self.stream)
        else:
            self.d_mask.zero_()
        self.dist.all_gather_into_tensor(self.d_gather, self.d_mask)
        torch.cuda.synchronize()
        good = torch.equal(got, self.d_gather) and self.ctx.stats()["watchdog"] == 0
        same = torch.tensor([int(good)], device=got.device)
        self.dist.all_reduce(same, op=self.dist.ReduceOp.MIN)
        if int(same.item()) == 0 and self.pg is not None:
            assert mode != "peer", "peer gather disagrees with the NCCL all-gather"
            self.pg, self.note = None, "all_gather of the validity bitmask (NCCL; the peer gather failed its check)"
            return "NCCL"
        assert int(same.item()) == 1
        return "gathered bitmask == NCCL all-gather on every rank"

    def launches_per_step(self):
        # kernels of ours: classify + descent (the gather rides in the descent's last CTA); an empty
        # shard still pushes its slot with one small kernel
        if self.n > 0:
            return 2
        return 1 if (self.pg is not None and self.world > 1) else 0


def timed_steps(torch, dist, world, dev, stream, flush, fn, steps):
    """K steps, CUDA events on the launching stream around each, L2 flushed before each outside the
    event pair, barrier + synchronize on both sides; returns the max over ranks of the summed time (ms)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for e0, e1 in evs:
        flush.fill_(1)
        e0.record(stream)
        fn()
        e1.record(stream)
    barrier()
    total = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    return float(total.item())


def main():
    args = parse()
    divert_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import robocheck_b200 as rc
    from robocheck_b200 import sharding as sh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=300))
    dev = torch.device("cuda", local)
    n = args.confs
    words = (n + 31) // 32

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = rc.Context(device=local)
    confs = rc.createAlphaBotConfigurations(n, True, 508 + rank)
    d_confs = torch.from_numpy(confs).to(dev)
    d_valid = torch.zeros(n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    # ---- headline: BASELINE configs[1], every rank its own 100,000 (weak scaling) -------------------
    head = Sharded(rc, dist, ctx, dev, d_confs, n, words, world, rank, args.gather, stream, d_valid)
    for _ in range(max(args.warmup, 3)):  # never fewer than three warm-up steps
        head.step()
    gather_check = head.verify_gather(args.gather)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # >= 1 s of the same steps, untimed, right before the timed ones: the device is at its working
    # clocks and the clock sampler (100 ms period) sees more than a handful of samples under load
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < args.preheat:
        for _ in range(50):
            head.step()
        torch.cuda.synchronize()
    total_ms = timed_steps(torch, dist, world, dev, stream, flush, head.step, args.steps)
    value = world * n * args.steps / (total_ms * 1e-3)
    launches = head.launches_per_step() * args.steps

    # ---- end to end: host buffers in, host flags out, through the public host call ---------------
    h_confs = torch.from_numpy(confs).pin_memory()
    h_valid = torch.zeros(n, dtype=torch.uint8).pin_memory()
    np_confs, np_valid = h_confs.numpy(), h_valid.numpy()
    lib = rc.lib()
    feed = os.environ.get("RC_FEED", "1")
    e2e_api = ("rc_check_host (pinned host buffers; single launch: the kernel reads the configurations from the "
               "caller's buffer over PCIe, 2.4 MB per step, and classifies them itself; flags copied back)"
               if feed != "0" else "rc_check_host (pinned host buffers; H2D copy, two launches, D2H copy)")
    h_mask_all = torch.zeros(world * words, dtype=torch.int32).pin_memory()
    d_stage = torch.empty(n * 6, dtype=torch.float32, device=dev)
    h_flat = h_confs.view(-1)

    if world > 1 and head.pg is not None:
        # N > 1: the call a rank of a sharded job makes -- host configurations in, the gathered bitmask
        # of ALL ranks back in host memory; the collective is inside the timed region
        e2e_api = ("sharding.PeerGather.check_host: H2D of the rank's configurations, rc_check_gather, D2H of "
                   "the gathered bitmask of all ranks")
        d2h_bytes = world * words * 4

        def e2e_step():
            head.pg.check_host(ctx, h_flat, d_stage, n, h_mask_all, stream)
    else:
        d2h_bytes = n

        def e2e_step():
            code = lib.rc_check_host(ctx._h, np_confs.ctypes.data, n, np_valid.ctypes.data, rc.RC_MODE_CULLED)
            assert code == 0

    for _ in range(3):
        e2e_step()
    # K synchronous steps, wall clock, max over ranks; repeated three times and the median kept: the
    # host-driven region is ~10 ms long and one nvidia-smi sample or scheduler hiccup inside it moves
    # the figure by 20 %
    e2e_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()   # synchronous: returns when the result is in host memory
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_runs.append(float(e2e_s.item()))
    e2e_value = world * n * args.steps / sorted(e2e_runs)[1]
    ctx.check(d_confs, n, d_valid, None, rc.RC_MODE_CULLED, stream)
    torch.cuda.synchronize()
    gpu_flags_now = d_valid.cpu().numpy()
    if world > 1 and head.pg is not None:
        mine = sh.unpack_bitmask(h_mask_all.numpy()[rank * words:(rank + 1) * words], words * 32)[:n]
        assert np.array_equal(mine, gpu_flags_now), "end-to-end gathered bitmask != device-resident flags"
    else:
        assert np.array_equal(np_valid, gpu_flags_now)
    # the same K host batches through the pipelined calls (planner loop: batch k+1 is submitted before the
    # flags of batch k are read, its copy overlaps the kernels of batch k); every step still moves its
    # inputs host->device and its flags device->host inside the timed region.  Reported next to e2e,
    # which stays the synchronous call.
    h_valid2 = torch.zeros(n, dtype=torch.uint8).pin_memory()
    outs = (np_valid, h_valid2.numpy())
    pipe_runs = []
    for _ in range(3):
        for o in outs:
            o.fill(7)
        barrier()
        t0 = time.perf_counter()
        prev = ctx.check_host_submit(np_confs, outs[0])
        for k in range(1, args.steps):
            cur = ctx.check_host_submit(np_confs, outs[k & 1])
            ctx.check_host_wait(prev)
            prev = cur
        ctx.check_host_wait(prev)
        barrier()
        pipe_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pipe_s, op=dist.ReduceOp.MAX)
        pipe_runs.append(float(pipe_s.item()))
    assert all(np.array_equal(o, gpu_flags_now) for o in outs[:min(2, args.steps)])
    e2e_pipelined = world * n * args.steps / sorted(pipe_runs)[1]
    clocks = sampler.stop() if rank == 0 else None

    # ---- north_star scaling workloads: ONE global set sharded over the ranks (strong scaling) -----
    scale = {}
    if not args.no_scale:
        # C4: BASELINE configs[3], 1 M configurations of the reference's sampler, seed 510
        n4 = args.c4
        lo, hi, per = sh.shard_bounds(n4, world, rank)
        all4 = rc.createAlphaBotConfigurations(n4, False, 510)
        d4 = torch.from_numpy(np.ascontiguousarray(all4[lo:hi])).to(dev) if hi > lo else torch.zeros(6, device=dev)
        run4 = Sharded(rc, dist, ctx, dev, d4, hi - lo, per // 32, world, rank, args.gather, stream)
        for _ in range(3):
            run4.step()
        chk4 = run4.verify_gather(args.gather)
        ms4 = timed_steps(torch, dist, world, dev, stream, flush, run4.step, args.scale_steps)
        bits4 = sh.unpack_bitmask(run4.step().cpu().numpy(), world * per)[:n4]
        scale["scale_c4"] = {
            "workload": f"BASELINE configs[3]: alpha1.0, {n4} configurations of the reference's sampler (seed 510), "
                        f"contiguous shards of {per}, bitmask gathered on every rank",
            "value": n4 * args.scale_steps / (ms4 * 1e-3), "unit": "configs/s", "ms": ms4 / args.scale_steps,
            "n_total": n4, "steps": args.scale_steps, "scaling": "strong", "collective": run4.note,
            "collective_check": chk4, "valid_fraction": float(bits4.mean()),
            "gpu_launches": run4.launches_per_step() * args.scale_steps}
        launches += run4.launches_per_step() * args.scale_steps
        del all4, d4, run4
        # C5: BASELINE configs[4], robot vs the obstacle subdivided five times (1,032,192 triangles),
        # 10 M configurations, nine blocks of 1,024 "easy" then one "hard" (SURVEY 8d), generated on
        # the device per shard: no 240 MB host-to-device copy anywhere
        n5 = args.c5
        rob_m, obs_m = rc.alpha_meshes()
        t0 = time.perf_counter()
        obs5 = obs_m.subdivide(args.c5_levels)
        ctx5 = rc.Context(rob_m, obs5, device=local)
        build5 = time.perf_counter() - t0
        lo, hi, per = sh.shard_bounds(n5, world, rank)
        cnt5 = hi - lo
        d5 = torch.empty(max(cnt5, 1) * 6, dtype=torch.float32, device=dev)
        ctx5.generate_device_mix(d5, lo, cnt5, rc.alpha_box(False), rc.alpha_box(True), 1024, 10, 511, stream)
        run5 = Sharded(rc, dist, ctx5, dev, d5, cnt5, per // 32, world, rank, args.gather, stream)
        run5.step()
        chk5 = run5.verify_gather(args.gather)
        ms5 = timed_steps(torch, dist, world, dev, stream, flush, run5.step, args.scale_steps)
        bits5 = sh.unpack_bitmask(run5.step().cpu().numpy(), world * per)[:n5]
        scale["scale_c5"] = {
            "workload": f"BASELINE configs[4]: alpha1.0 robot vs obstacle subdivided x{args.c5_levels} "
                        f"({obs5.n_t} triangles, tables in global memory), {n5} configurations generated on the "
                        "device per shard (blocks of 1,024: nine easy, one hard; seed 511), bitmask gathered",
            "value": n5 * args.scale_steps / (ms5 * 1e-3), "unit": "configs/s", "ms": ms5 / args.scale_steps,
            "n_total": n5, "steps": args.scale_steps, "scaling": "strong", "collective": run5.note,
            "collective_check": chk5, "valid_fraction": float(bits5.mean()),
            "context_build_s": build5, "gpu_launches": run5.launches_per_step() * args.scale_steps}
        launches += run5.launches_per_step() * args.scale_steps
        watchdog5 = ctx5.stats()["watchdog"]
        del run5, d5
        ctx5.close()
        assert watchdog5 == 0

    if world > 1:
        # every rank leaves the group here, after the last collective of the timed regions
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- explanations (rank 0, untimed w.r.t. the numbers above) ------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    fp32_peak = ctx.fp32_peak_tflops()
    d_mask = torch.zeros(words, dtype=torch.int32, device=dev)
    ctx.stats_enable(True)
    ctx.stats(reset=True)
    # rank 0 only from here on: the kernel alone, never the collective
    ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, stream); torch.cuda.synchronize()
    st = ctx.stats(reset=True)
    ctx.stats_enable(False)
    step_s = total_ms * 1e-3 / args.steps
    # EXECUTED work of one launch, from the kernel's own counters (statistics build, same workload):
    # box tests, robot triangles set up in leaf rounds, triangle-box tests, plane tests, finished pairs
    exe_flop = executed_flop(st)
    # algorithmic work of the reference formulation: 19,008 flop per configuration for
    # transform + AABB, 18 flop per decided triangle pair; a narrow-phase configuration found valid
    # decides all 1008 x 1008 pairs, a colliding one is counted with a single pair
    decided = st["narrow_valid"] * PAIRS_PER_CONFIG + st["narrow_invalid"]
    alg_flop = 19008.0 * n + 18.0 * decided
    # hardware view of the same kernels from the committed ncu capture (not live; profiles/)
    prof, prof_name = {}, None
    for name in ("ncu_r02_metrics.json", "ncu_r01_metrics.json"):
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", name)))
            prof_name = name
            break
        except (OSError, ValueError):
            continue

    def hw_view(kernel):
        p = prof.get(kernel)
        if not p:
            return None
        return {"source": f"profiles/{prof_name} (ncu --set full, same kernel, not live)",
                "issue_slots_busy_pct": p.get("issue_active_pct"), "fma_pipe_pct": p.get("fma_pipe_pct"),
                "alu_pipe_pct": p.get("alu_pipe_pct"), "dram_throughput_pct": p.get("dram_throughput_pct"),
                "lanes_per_instruction": p.get("lanes_per_instruction"),
                "warp_instructions": p.get("inst_executed"),
                "shared_bank_conflict_share": p.get("bank_conflict_share")}

    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    roofline = {
        "kernel": "check_culled_kernel", "bound": "fp32",
        "achieved": exe_flop / step_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": exe_flop / step_s / 1e12 / fp32_peak, "traffic": None,
        "peak_source": "FP32 FMA rate measured live by rc_measure_fp32_peak (MEASURED_PEAKS.json "
                       "has no FP32 entry)",
        "note": "EXECUTED flops of one launch (device counters of the statistics build x flop per event, "
                "bench.py FLOP_*) over the event-timed step; the kernel is bound by instruction issue on "
                "non-FP work (queues, votes, shared-memory gathers), see hw and issue",
        "executed": {k: st[k] for k in ("node_pairs", "leaf_pairs", "tri_pairs", "prefilter_pairs",
                                        "plane1_pairs", "tail_pairs", "overflow")},
        "executed_gflop": exe_flop / 1e9,
        "work_avoided_x": alg_flop / max(exe_flop, 1.0),
        "algorithmic_tflops_reference_formulation": alg_flop / step_s / 1e12,
        "mix": {k: st[k] for k in ("broad_cleared", "sphere_cleared", "narrow_valid", "narrow_invalid")},
    }
    if prof.get("check_culled_kernel"):
        pk = prof["check_culled_kernel"]
        roofline["traffic"] = pk.get("dram_bytes")
        roofline["traffic_unit"] = "bytes per launch (ncu dram__bytes_read+write, 100,000 configurations)"
        roofline["hw"] = hw_view("check_culled_kernel")
        if pk.get("inst_executed"):
            sms = peaks.get("sm_count", 148)
            roofline["issue"] = {
                "warp_instructions_per_launch": pk["inst_executed"],
                "frac_of_issue_slots": pk["inst_executed"] / (sms * 4 * sm_mhz * 1e6 * step_s),
                "note": "ncu smsp__inst_executed.sum of the committed capture / (SMs x 4 schedulers x SM clock "
                        "x live step time)"}
    out = {
        "metric": "collision-checked configs/sec", "value": value, "unit": "configs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "configs_per_gpu": n, "mode": "culled fused kernel",
                   "l2": "flushed before every timed step (256 MiB fill, outside the event pair)",
                   "preheat_s": args.preheat,
                   "collective": head.note, "collective_check": gather_check},
        "e2e": {"value": e2e_value, "unit": "configs/s", "h2d_bytes_per_step": n * 24,
                "d2h_bytes_per_step": d2h_bytes, "api": e2e_api,
                "timing": "median of 3 runs of K synchronous steps each (wall clock, max over ranks)",
                "runs_configs_per_s": [world * n * args.steps / t for t in e2e_runs],
                "pipelined": {"value": e2e_pipelined, "unit": "configs/s",
                              "api": "rc_check_host_submit / rc_check_host_wait, two batches in flight",
                              "note": "same K batches and bytes per step; the copy of batch k+1 overlaps the "
                                      "kernels of batch k (not the headline: e2e.value is the synchronous call)"}},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "valid_fraction": float(d_valid.float().mean().item()),
    }
    out.update(scale)

    if not args.no_extras and world == 1:
        # un-culled kernel on a bounded sample: the FP32-bound formulation of SURVEY 8(d)
        nb = min(n, 32768)
        ctx.stats_enable(True); ctx.stats(reset=True)
        ctx.check(d_confs, nb, d_valid, None, rc.RC_MODE_BRUTE, stream); torch.cuda.synchronize()
        sb = ctx.stats(reset=True); ctx.stats_enable(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        e0.record(stream)
        ctx.check(d_confs, nb, d_valid, None, rc.RC_MODE_BRUTE, stream)
        e1.record(stream); torch.cuda.synchronize()
        tb = e0.elapsed_time(e1) * 1e-3
        flop_b = 19008.0 * nb + 18.0 * (sb["narrow_valid"] * PAIRS_PER_CONFIG + sb["narrow_invalid"])
        out["roofline_brute"] = {
            "kernel": "check_brute_kernel", "bound": "fp32", "achieved": flop_b / tb / 1e12,
            "peak": fp32_peak, "unit": "TFLOP/s", "frac": flop_b / tb / 1e12 / fp32_peak,
            "configs": nb, "ms": tb * 1e3, "configs_per_s": nb / tb, "hw": hw_view("check_brute_kernel")}
        # materialising transform+AABB+broad stage: HBM-bound, 9,553 B per configuration
        ns = min(n, 100_000)
        nv = ctx.n_rob_v
        dx, dy, dz = (torch.empty(ns * nv, dtype=torch.float32, device=dev) for _ in range(3))
        d_box = torch.empty(ns * 6, dtype=torch.float32, device=dev)
        ts = []
        for it in range(6):
            # no L2 flush here: one launch writes 955 MB, far more than the 126 MB L2 (a write-fill flush
            # would leave the L2 full of dirty lines whose write-back is then charged to this kernel)
            e0.record(stream)
            ctx.broad_fused(d_confs, ns, dx, dy, dz, d_box, d_valid, stream)
            e1.record(stream); torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1) * 1e-3)
        t_stage = sum(ts) / len(ts)
        stage_bytes = 9553.0 * ns
        out["roofline_stage"] = {
            "kernel": "transform_soa_kernel (broadPhaseFusedKernel_sep stage)", "bound": "hbm",
            "achieved": stage_bytes / t_stage / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": stage_bytes / t_stage / 1e9 / hbm_peak, "peak_source": hbm_src,
            "configs": ns, "ms": t_stage * 1e3, "configs_per_s": ns / t_stage,
            "l2": "output of one launch (955 MB) exceeds L2; not flushed",
            "traffic": (prof.get("transform_soa_kernel") or {}).get("dram_bytes"),
            "hw": hw_view("transform_soa_kernel")}
        del dx, dy, dz
        from oracle import pyoracle as po
        out["cpu_baseline"], _, _ = cpu_reference_throughput(rc, po, confs, seconds=12.0)
        # second same-box comparator: tree-culled CPU checker with the reference's per-pair
        # arithmetic, the stand-in for FCL's BVH path (FCL is not installable here); it finishes the
        # whole workload, so it also checks every flag of the timed steps
        rob_m, obs_m = rc.alpha_meshes()
        orob, oobs = po.OracleMesh(*rob_m.arrays()), po.OracleMesh(*obs_m.arrays())
        threads = os.cpu_count() or 1
        po.check_bvh(orob, oobs, confs[:2048], po.VARIANT_GPU, threads)  # warm-up
        t0 = time.perf_counter()
        cpu_flags = po.check_bvh(orob, oobs, confs, po.VARIANT_GPU, threads)
        dt_all = time.perf_counter() - t0
        n1 = min(n, 20000)
        t0 = time.perf_counter()
        po.check_bvh(orob, oobs, confs[:n1], po.VARIANT_GPU, 1)
        dt_one = time.perf_counter() - t0
        out["cpu_baseline_tree"] = {
            "value": n / dt_all, "unit": "configs/s", "cores": threads, "kind": "port",
            "sample": f"all {n} workload configurations, {dt_all:.2f} s; oracle/oracle_bvh.c: dual-tree "
                      "descent + the reference's per-pair arithmetic (stand-in for FCL's BVH path)",
            "single_core_value": n1 / dt_one,
            "single_core_sample": f"first {n1} configurations, {dt_one:.2f} s"}
        out["vs_cpu_tree"] = {"device_resident_x": value / (n / dt_all), "e2e_x": e2e_value / (n / dt_all),
                              "note": "value and e2e.value over cpu_baseline_tree.value (same box, all host cores)"}
        out["vs_cpu_reference_brute_force"] = {
            "device_resident_x": value / out["cpu_baseline"]["value"], "e2e_x": e2e_value / out["cpu_baseline"]["value"],
            "note": "over cpu_baseline.value: the reference's narrowPhaseBaseline, an exhaustive sweep -- a weak "
                    "comparator, quoted because it is the reference's own CPU code"}
        out["fcl"] = {
            "available": False,
            "reason": "FCL, libccd and Eigen are not installed in this image and cannot be installed (no network); "
                      "the 100 k and 1 M sets therefore carry no FCL verdicts, only the two shipped 10 k sets do "
                      "(frozen tags, 0 disagreements of 20,000 in tests/)",
            "published_other_hardware": {"value": 206600.0, "unit": "configs/s", "what": "FCL -O3, hard 100 k, one "
                                         "core of an i7-8700K (BASELINE.md)"},
            "vs_published_x": {"device_resident": value / 206600.0, "e2e": e2e_value / 206600.0},
            "stand_ins_same_box": ["cpu_baseline (reference brute force, all cores)",
                                   "cpu_baseline_tree (tree-culled CPU checker, all cores)"]}
        # parity of the timed workload against checkers that do NOT share the kernel's culling:
        # the exhaustive kernel, and the unmodified reference pipeline run on this GPU
        ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, stream); torch.cuda.synchronize()
        gpu_flags = d_valid.cpu().numpy()
        ctx.check(d_confs, n, d_valid, None, rc.RC_MODE_BRUTE, stream); torch.cuda.synchronize()
        brute_flags = d_valid.cpu().numpy()
        parity = {"checker": "oracle/oracle_bvh.c (CPU, tree-culled), every configuration of the workload",
                  "configs": n, "flag_mismatches": int((gpu_flags != cpu_flags).sum()),
                  "e2e_flag_mismatches": int((np_valid != cpu_flags).sum()),
                  "brute_flag_mismatches": int((gpu_flags != brute_flags).sum()),
                  "brute_checker": "RC_MODE_BRUTE: every one of the 1008 x 1008 pairs of every configuration"}
        d_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
        nf, nc, _ = ctx.audit(d_confs, n, cap=0)
        parity["audit"] = {"flags_differ": nf, "hit_pair_counts_differ": nc,
                           "what": "rc_audit: culled vs exhaustive, flags and per-configuration hitting-pair counts"}
        del d_cnt
        if po.ref_available():
            try:
                from tests import refpipe
                t0 = time.perf_counter()
                ref_flags = refpipe.oracle_gpu_flags(rc, po, confs)
                parity["reference_gpu_flag_mismatches"] = int((gpu_flags != ref_flags).sum())
                parity["reference_gpu_checker"] = ("broadPhaseFused_sep of the unmodified reference (oracle/_ref), "
                                                   f"run on this GPU: {time.perf_counter() - t0:.1f} s")
            except Exception as exc:  # the checker is optional evidence, never the product
                parity["reference_gpu_checker"] = f"failed: {type(exc).__name__}: {exc}"
        out["parity"] = parity

    emit(out)


if __name__ == "__main__":
    main()
