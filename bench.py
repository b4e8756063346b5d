#!/usr/bin/env python
"""bench.py -- collision-checked configurations per second on the alpha1.0 puzzle.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference CPU path)

Workload (config.workload): BASELINE.json configs[1], "alpha1.0 hard_confs100,000": 100,000
near-contact configurations drawn like the reference's generator (Utils_rai.cu:111-136, 322-348;
the shipped 100k file is a missing blob, so the set is regenerated with seed 508 + rank).  A
"step" is one pass of the hot path over the batch.  With N GPUs every rank checks its own 100,000
configurations (weak scaling) and the per-configuration validity bitmask is all-gathered over NCCL
inside the timed region; value = all configurations of all ranks / max-over-ranks device time.

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


def main():
    args = parse()
    divert_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import robocheck_b200 as rc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=180))
    dev = torch.device("cuda", local)
    n = args.confs
    words = (n + 31) // 32

    ctx = rc.Context(device=local)
    confs = rc.createAlphaBotConfigurations(n, True, 508 + rank)
    d_confs = torch.from_numpy(confs).to(dev)
    d_valid = torch.zeros(n, dtype=torch.uint8, device=dev)
    # every rank's kernel writes its words straight into its slot of the gather buffer
    d_gather = torch.zeros(world * words, dtype=torch.int32, device=dev)
    d_mask = d_gather[rank * words:(rank + 1) * words]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    # N>1: the all-gather of the bitmask fused behind the check over peer-mapped memory (NVLink),
    # falling back to the NCCL all-gather when symmetric memory is not available on the box
    pg, gather_note = None, "none"
    if world > 1:
        gather_note = "all_gather of the validity bitmask (NCCL)"
        if args.gather != "nccl":
            try:
                from robocheck_b200 import sharding as sh
                pg = sh.PeerGather(rc, dist, dev, words)
            except Exception as exc:
                pg = None
                print(f"[bench] rank {rank}: peer gather unavailable ({type(exc).__name__}: {exc})", file=sys.stderr)
            okf = torch.tensor([1 if pg is not None else 0], device=dev)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if int(okf.item()) == 0:
                assert args.gather != "peer", "--gather peer: symmetric memory not available"
                pg = None
            else:
                gather_note = ("bitmask words stored into the peers' buffers over NVLink by a kernel behind the "
                               "check (rc_check_gather, torch symmetric memory), no communicator call")

    def step():
        if pg is not None:
            return pg.check(ctx, d_confs, n, d_valid, stream)
        ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, stream)
        if world > 1:
            dist.all_gather_into_tensor(d_gather, d_mask)
        return d_gather

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):  # never fewer than three warm-up steps
        step()
    # untimed, before the timed steps: the gathered bitmask of one step against an NCCL all-gather of the
    # same words; a peer gather that disagrees (or whose wait gave up) is not used for the timed steps
    gather_check = None
    if world > 1:
        got = step().clone()
        ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, stream)
        dist.all_gather_into_tensor(d_gather, d_mask)
        torch.cuda.synchronize()
        good = torch.equal(got, d_gather) and ctx.stats()["watchdog"] == 0
        same = torch.tensor([int(good)], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if int(same.item()) == 0 and pg is not None:
            assert args.gather != "peer", "peer gather disagrees with the NCCL all-gather"
            pg, gather_note = None, "all_gather of the validity bitmask (NCCL; the peer gather failed its check)"
            gather_check = "NCCL"
        else:
            assert int(same.item()) == 1
            gather_check = "gathered bitmask == NCCL all-gather on every rank"
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:
        flush.fill_(1)          # L2 flush, outside the event pair
        e0.record(stream)
        step()
        e1.record(stream)
    barrier()
    ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * n * args.steps / (total_ms * 1e-3)
    # kernels of ours per step: classify + descent (+ the gather push at N>1); memsets and NCCL are not ours
    launches = (3 if pg is not None else 2) * args.steps
    # ---- end to end: host buffers in, host flags out, through the public host call ---------------
    h_confs = torch.from_numpy(confs).pin_memory()
    h_valid = torch.zeros(n, dtype=torch.uint8).pin_memory()
    np_confs, np_valid = h_confs.numpy(), h_valid.numpy()
    lib = rc.lib()

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
            e2e_step()   # synchronous: returns when the flags are in host memory
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_runs.append(float(e2e_s.item()))
    e2e_value = world * n * args.steps / sorted(e2e_runs)[1]
    assert np.array_equal(np_valid, d_valid.cpu().numpy())
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
    gpu_flags_now = d_valid.cpu().numpy()
    assert all(np.array_equal(o, gpu_flags_now) for o in outs[:min(2, args.steps)])
    e2e_pipelined = world * n * args.steps / sorted(pipe_runs)[1]
    clocks = sampler.stop() if rank == 0 else None

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
    ctx.stats_enable(True)
    ctx.stats(reset=True)
    # rank 0 only from here on: the kernel alone, never the collective
    ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, stream); torch.cuda.synchronize()
    st = ctx.stats(reset=True)
    ctx.stats_enable(False)
    step_s = total_ms * 1e-3 / args.steps
    # algorithmic work of one launch: 19,008 flop per configuration for transform +
    # AABB, 18 flop per decided triangle pair; a narrow-phase configuration found valid decides
    # all 1008 x 1008 pairs, a colliding one is counted with a single pair.
    decided = st["narrow_valid"] * PAIRS_PER_CONFIG + st["narrow_invalid"]
    alg_flop = 19008.0 * n + 18.0 * decided
    roofline = {
        "kernel": "check_culled_kernel", "bound": "fp32",
        "achieved": alg_flop / step_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": alg_flop / step_s / 1e12 / fp32_peak, "traffic": None,
        "peak_source": "FP32 FMA rate measured live by rc_measure_fp32_peak (MEASURED_PEAKS.json "
                       "has no FP32 entry)",
        "note": "algorithmic flops of the reference formulation (18 flop x every decided pair); "
                "BVH culling decides most pairs without touching them, so frac > 1 is expected - "
                "see roofline_brute for the un-culled kernel and profiles/ for the ncu pipe view",
        "executed": {k: st[k] for k in ("node_pairs", "leaf_pairs", "tri_pairs", "overflow")},
        "mix": {k: st[k] for k in ("broad_cleared", "sphere_cleared", "narrow_valid", "narrow_invalid")},
    }
    # hardware view of the same kernels from the committed ncu capture (not live; profiles/)
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_r01_metrics.json")))
    except (OSError, ValueError):
        pass

    def hw_view(kernel):
        p = prof.get(kernel)
        if not p:
            return None
        return {"source": "profiles/ncu_r01_summary.md (ncu --set full, same kernel, not live)",
                "issue_slots_busy_pct": p.get("issue_active_pct"), "fma_pipe_pct": p.get("fma_pipe_pct"),
                "alu_pipe_pct": p.get("alu_pipe_pct"), "dram_throughput_pct": p.get("dram_throughput_pct"),
                "lanes_per_instruction": p.get("lanes_per_instruction")}

    if prof.get("check_culled_kernel"):
        roofline["traffic"] = prof["check_culled_kernel"].get("dram_bytes")
        roofline["traffic_unit"] = "bytes per launch (ncu dram__bytes_read+write, 100,000 configurations)"
        roofline["hw"] = hw_view("check_culled_kernel")
    out = {
        "metric": "collision-checked configs/sec", "value": value, "unit": "configs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "configs_per_gpu": n, "mode": "culled fused kernel",
                   "l2": "flushed before every timed step (256 MiB fill, outside the event pair)",
                   "collective": gather_note, "collective_check": gather_check},
        "e2e": {"value": e2e_value, "unit": "configs/s", "h2d_bytes_per_step": n * 24,
                "d2h_bytes_per_step": n, "api": "rc_check_host (pinned host buffers)",
                "timing": "median of 3 runs of K synchronous steps each (wall clock, max over ranks)",
                "runs_configs_per_s": [world * n * args.steps / t for t in e2e_runs],
                "pipelined": {"value": e2e_pipelined, "unit": "configs/s",
                              "api": "rc_check_host_submit / rc_check_host_wait, two batches in flight",
                              "note": "same K batches and bytes per step; the copy of batch k+1 overlaps the "
                                      "kernels of batch k (not the headline: e2e.value is the synchronous call)"}},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "valid_fraction": float(d_valid.float().mean().item()),
    }

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
        ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, stream); torch.cuda.synchronize()
        gpu_flags = d_valid.cpu().numpy()
        out["parity"] = {"checker": "oracle/oracle_bvh.c (CPU), every configuration of the workload",
                         "configs": n, "flag_mismatches": int((gpu_flags != cpu_flags).sum()),
                         "e2e_flag_mismatches": int((np_valid != cpu_flags).sum())}

    emit(out)


if __name__ == "__main__":
    main()
