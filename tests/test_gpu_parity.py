"""Parity tests proper: the CUDA path, called through the C ABI, against (a) the CPU oracle,
(b) the committed golden fixtures (FCL tags, reference CPU flags), (c) ORACLE-GPU = the
reference's own kernels compiled for sm_100a (oracle/_ref/libref.so) on the same inputs.
Integer/flag work: bit-exact.  Transformed vertices: bit-exact vs the reference kernel and the
oracle (tolerance stated by north_star is 1e-6 relative; we assert equality)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def run_check(ctx, confs, mode, bitmask=False):
    torch = _torch()
    d_confs = torch.from_numpy(np.ascontiguousarray(confs, dtype=np.float32)).cuda()
    n = len(confs)
    d_valid = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    d_mask = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda") if bitmask else None
    ctx.check(d_confs, n, d_valid, d_mask, mode, torch.cuda.current_stream())
    torch.cuda.synchronize()
    if bitmask:
        bits = np.unpackbits(d_mask.cpu().numpy().view(np.uint8), bitorder="little")[:n]
        return d_valid.cpu().numpy(), bits
    return d_valid.cpu().numpy()


@pytest.fixture(scope="module")
def mixed_confs(rc):
    return np.concatenate([rc.createAlphaBotConfigurations(1500, True, 508),
                           rc.createAlphaBotConfigurations(500, False, 509)]).astype(np.float32)


@pytest.fixture(scope="module")
def mixed_oracle(po, omeshes, mixed_confs):
    return po.check(omeshes[0], omeshes[1], mixed_confs, po.VARIANT_GPU)


def test_brute_matches_oracle(rc, ctx, mixed_confs, mixed_oracle):
    got = run_check(ctx, mixed_confs, rc.RC_MODE_BRUTE)
    assert np.array_equal(got, mixed_oracle)


def test_culled_matches_oracle(rc, ctx, mixed_confs, mixed_oracle):
    got, bits = run_check(ctx, mixed_confs, rc.RC_MODE_CULLED, bitmask=True)
    assert np.array_equal(got, mixed_oracle)
    assert np.array_equal(bits, mixed_oracle)


def test_golden_fcl_tags(rc, ctx, golden):
    """Every FCL verdict of the reference's shipped sets is reproduced (the expected
    disagreement count is 0, so the contact tolerance is never needed)."""
    for key in ("easy_confs10k", "hard_confs10k"):
        g = golden[key]
        got = run_check(ctx, g["confs"], rc.RC_MODE_CULLED)
        assert int((got != g["tags"]).sum()) == 0, key


def test_golden_reference_cpu_flags(rc, ctx, golden):
    g = golden["refcpu_hard"]
    for mode in (rc.RC_MODE_CULLED, rc.RC_MODE_BRUTE):
        got = run_check(ctx, g["confs"], mode)
        assert np.array_equal(got, g["flags"])


def test_check_host_matches_device(rc, ctx, mixed_confs, mixed_oracle):
    assert np.array_equal(ctx.check_host(mixed_confs), mixed_oracle)
    assert np.array_equal(ctx.check_host(mixed_confs, rc.RC_MODE_BRUTE), mixed_oracle)


def test_empty_and_ragged_sizes(rc, ctx, po, omeshes):
    torch = _torch()
    ctx.check(None, 0, torch.zeros(1, dtype=torch.uint8, device="cuda"), None)
    assert len(ctx.check_host(np.zeros((0, 6), np.float32))) == 0
    confs = rc.createAlphaBotConfigurations(97, True, 7)
    want = po.check(omeshes[0], omeshes[1], confs, po.VARIANT_GPU)
    for n in (1, 31, 32, 33, 97):
        got, bits = run_check(ctx, confs[:n], rc.RC_MODE_CULLED, bitmask=True)
        assert np.array_equal(got, want[:n])
        assert np.array_equal(bits, want[:n])


def test_culled_equals_brute_large(rc, ctx):
    """Size-independent property at a scale the CPU oracle cannot reach: the culled kernel and
    the exhaustive kernel agree on every flag (culling never drops a hitting pair)."""
    confs = np.concatenate([rc.createAlphaBotConfigurations(60000, True, 4242),
                            rc.createAlphaBotConfigurations(20000, False, 4243)])
    a = run_check(ctx, confs, rc.RC_MODE_CULLED)
    b = run_check(ctx, confs, rc.RC_MODE_BRUTE)
    assert int((a != b).sum()) == 0
    st_on = ctx.stats_enable(True)
    ctx.stats(reset=True)
    run_check(ctx, confs, rc.RC_MODE_CULLED)
    st = ctx.stats(reset=True)
    ctx.stats_enable(False)
    assert st["overflow"] == 0
    assert st["broad_cleared"] + st["sphere_cleared"] + st["narrow_valid"] + st["narrow_invalid"] == len(confs)
    assert st["narrow_valid"] + st["broad_cleared"] + st["sphere_cleared"] == int(a.sum())


@pytest.mark.parametrize("knobs", [
    {},                                                       # defaults
    {"RC_DONATE_MIN": "40", "RC_MAIL_SLOTS": "16"},           # hand over early and often
    {"RC_TAIL_THRESH": "8", "RC_LEAF_THRESH": "4"},           # many short finishing rounds
    {"RC_DONATE_MIN": "200", "RC_MAIL_SLOTS": "1"},           # one slot, rare hand-overs
    {"RC_SHARE": "0"},                                        # no hand-over at all
    {"RC_BOUND": "0"},                                        # one cost class, no vertex-group bound
    {"RC_PDL": "0"},                                          # plain stream order instead of the programmatic dependent launch
    {"RC_CULLED_WARPS": "16"},
    {"RC_CULLED_WARPS": "24"},
])
def test_descent_scheduling_variants_equal_brute(rc, meshes, knobs, monkeypatch):
    """How the descent is scheduled (cost classes, hand-over of pieces between the warps of a CTA,
    queue thresholds) must never change a flag: 100,000 hard configurations, repeated three times
    (the hand-over is timing dependent), against the exhaustive kernel; the watchdog stays silent."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    confs = rc.createAlphaBotConfigurations(100000, True, 777)
    c = rc.Context(meshes[0], meshes[1], device=0)
    try:
        want = run_check(c, confs, rc.RC_MODE_BRUTE)
        c.stats_enable(True)
        c.stats(reset=True)
        got_stats = run_check(c, confs, rc.RC_MODE_CULLED)
        st = c.stats(reset=True)
        c.stats_enable(False)
        assert int((got_stats != want).sum()) == 0
        for _ in range(3):
            got = run_check(c, confs, rc.RC_MODE_CULLED)
            assert int((got != want).sum()) == 0
    finally:
        c.close()
    assert st["watchdog"] == 0 and st["overflow"] == 0
    if knobs.get("RC_SHARE") == "0":
        assert st["shared_pieces"] == 0
    assert st["broad_cleared"] + st["sphere_cleared"] + st["narrow_valid"] + st["narrow_invalid"] == len(confs)


@pytest.mark.parametrize("n_ctx,n", [(1, 5000), (2, 40001), (3, 64), (4, 100000), (4, 70)])
def test_check_multi_single_process(rc, ctx, meshes, n_ctx, n):
    """rc_check_multi: contiguous shards of whole bitmask words, every context checks its
    shard, the host gets the full bitmask and flags, and every device buffer the full bitmask (peer
    copies).  Several contexts on one device stand in for several devices here; the same call with
    one context per GPU runs in tests/test_gpu_multi.py."""
    torch = _torch()
    confs = np.concatenate([rc.createAlphaBotConfigurations(n - n // 4, True, 91),
                            rc.createAlphaBotConfigurations(n // 4, False, 92)])
    confs = confs[np.random.default_rng(5).permutation(n)]
    want = run_check(ctx, confs, rc.RC_MODE_CULLED)
    ctxs = [rc.Context(meshes[0], meshes[1], device=0) for _ in range(n_ctx)]
    try:
        per = rc.lib().rc_shard_size(n, n_ctx)
        assert per % 32 == 0 and per * n_ctx >= n
        words = (n + 31) // 32
        d_all = [torch.full((words,), -1, dtype=torch.int32, device="cuda") for _ in range(n_ctx)]
        torch.cuda.synchronize()
        mask, flags = rc.check_multi(ctxs, confs, want_flags=True, d_bitmask_all=d_all)
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:n]
        assert np.array_equal(flags, want)
        assert np.array_equal(bits, want)
        for t in d_all:
            got = np.unpackbits(t.cpu().numpy().view(np.uint8), bitorder="little")[:n]
            assert np.array_equal(got, want)
        # bitmask only, no device buffers
        mask2 = rc.check_multi(ctxs, confs)
        assert np.array_equal(mask2, mask)
    finally:
        for c in ctxs:
            c.close()


# ---- stage API -------------------------------------------------------------------------------
def test_transform_bit_exact_vs_oracle(rc, ctx, po, omeshes):
    torch = _torch()
    confs = np.concatenate([rc.createAlphaBotConfigurations(300, True, 11),
                            rc.createAlphaBotConfigurations(300, False, 12)]).astype(np.float32)
    n, nv = len(confs), ctx.n_rob_v
    d_confs = torch.from_numpy(confs).cuda()
    dx, dy, dz = (torch.zeros(n * nv, dtype=torch.float32, device="cuda") for _ in range(3))
    d_box = torch.zeros(n * 6, dtype=torch.float32, device="cuda")
    d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ctx.broad_fused(d_confs, n, dx, dy, dz, d_box, d_valid, torch.cuda.current_stream())
    torch.cuda.synchronize()
    ox, oy, oz = po.transform(confs, omeshes[0].x, omeshes[0].y, omeshes[0].z)
    for got, want in ((dx, ox), (dy, oy), (dz, oz)):
        got = got.cpu().numpy().reshape(n, nv)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    boxes = d_box.cpu().numpy().reshape(n, 6)
    obs_box = ctx.obstacle_aabb()
    assert np.array_equal(obs_box, po.aabb(omeshes[1].x, omeshes[1].y, omeshes[1].z))
    for i in range(n):
        want = po.aabb(ox[i], oy[i], oz[i])
        assert np.array_equal(boxes[i], want)
        assert int(d_valid[i]) == po.broad_valid(want, obs_box)
    # AoS stage (genTransformedCopies layout) carries the same bits
    d_aos = torch.zeros(n * nv * 3, dtype=torch.float32, device="cuda")
    ctx.transform_aos(d_confs, n, d_aos, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    aos = d_aos.cpu().numpy().reshape(n, nv, 3)
    assert np.array_equal(aos[..., 0], ox) and np.array_equal(aos[..., 1], oy) and np.array_equal(aos[..., 2], oz)
    # AABB stages on the materialised vertices, exact equality like test-generate-AABB.cu:83-96
    d_box2 = torch.zeros(n * 6, dtype=torch.float32, device="cuda")
    ctx.generate_aabb(d_aos, nv, n, d_box2, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(d_box2.cpu().numpy().reshape(n, 6), boxes)
    d_box3 = torch.zeros(n * 6, dtype=torch.float32, device="cuda")
    ctx.generate_aabb_soa(dx, dy, dz, nv, n, d_box3, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(d_box3.cpu().numpy().reshape(n, 6), boxes)
    assert np.array_equal(ctx.generate_aabb_host(aos, nv, n), boxes)
    # broad-phase stage
    d_valid2 = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ctx.broad_phase(n, d_box, d_valid2, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(d_valid, d_valid2)


def test_stage_pipeline_equals_fused(rc, ctx, po, omeshes):
    """transform+AABB+broad stage followed by the narrow stage == fused entry point == oracle."""
    torch = _torch()
    confs = rc.createAlphaBotConfigurations(600, True, 13)
    n, nv = len(confs), ctx.n_rob_v
    d_confs = torch.from_numpy(confs).cuda()
    dx, dy, dz = (torch.zeros(n * nv, dtype=torch.float32, device="cuda") for _ in range(3))
    d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream()
    ctx.broad_fused(d_confs, n, dx, dy, dz, None, d_valid, s)
    ctx.narrow_phase(dx, dy, dz, n, d_valid, s)
    torch.cuda.synchronize()
    want = po.check(omeshes[0], omeshes[1], confs, po.VARIANT_GPU)
    assert np.array_equal(d_valid.cpu().numpy(), want)
    # in/out semantics: entries already 1 are left untouched (narrow-phase.cu:578)
    d_valid.fill_(1)
    ctx.narrow_phase(dx, dy, dz, n, d_valid, s)
    torch.cuda.synchronize()
    assert int(d_valid.sum()) == n


def test_narrow_phase_host_tetrahedra(rc):
    """test-narrow-phase.cu:628-686 (test_baseline) through narrowPhase(): two configurations of a
    tetrahedron against a tetrahedron, the first interpenetrating, the second 100 units away
    -> res = [false, true]; and :688-739 (test_single_triangle): must stay valid."""
    rob = np.array([[0, 0, 0], [0, 1, 2], [0, 2, 0], [2, 2, 2],
                    [100, 0, 0], [100, 1, 2], [100, 2, 0], [102, 2, 2]], np.float32).reshape(2, 4, 3)
    obs = np.array([[1, 0.1, 0], [1.5, 1, 2.2], [1, 2.4, 0], [3.3, 2, 2]], np.float32)
    tris = np.array([[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]], np.int32)
    valid = np.zeros(2, np.uint8)
    rc.narrowPhase(tris, rob, tris, obs, valid)
    assert valid.tolist() == [0, 1]
    r = np.array([[1.441547, -14.800514, 62.841087], [-4.215309, 8.199282, 23.057938],
                  [1.883977, -15.487457, 62.381035]], np.float32).reshape(1, 3, 3)
    o = np.array([[1.681669, 2.616245, 1.069425], [3.561536, 0.677467, 1.707230],
                  [1.172210, 2.534812, 1.852433]], np.float32)
    one = np.zeros(1, np.uint8)
    rc.narrowPhase(np.array([[0, 1, 2]], np.int32), r, np.array([[0, 1, 2]], np.int32), o, one)
    assert one.tolist() == [1]
    # test_gpu (:742-836): 4096 configurations alternating collide / valid
    n = 4096
    alt = np.stack([rob[i % 2] for i in range(n)])
    flags = np.zeros(n, np.uint8)
    rc.narrowPhase(tris, alt, tris, obs, flags)
    assert np.array_equal(flags, np.arange(n) % 2)
    flags[:] = 1  # "all pre-valid -> untouched"
    rc.narrowPhase(tris, alt, tris, obs, flags)
    assert int(flags.sum()) == n


# ---- ORACLE-GPU: the reference's own CUDA kernels on this GPU --------------------------------
def test_flags_match_reference_gpu_pipeline(rc, ctx, po):
    if not po.ref_available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    from tests import refpipe
    confs = np.concatenate([rc.createAlphaBotConfigurations(20000, True, 601),
                            rc.createAlphaBotConfigurations(5000, False, 602)]).astype(np.float32)
    want = refpipe.oracle_gpu_flags(rc, po, confs)
    for mode in (rc.RC_MODE_CULLED, rc.RC_MODE_BRUTE):
        got = run_check(ctx, confs, mode)
        assert int((got != want).sum()) == 0


def test_vertices_match_reference_gpu_kernel(rc, ctx, po, omeshes):
    if not po.ref_available():
        pytest.skip("oracle/_ref/libref.so not built")
    torch = _torch()
    confs = np.concatenate([rc.createAlphaBotConfigurations(2000, True, 603),
                            rc.createAlphaBotConfigurations(2000, False, 604)]).astype(np.float32)
    rx, ry, rz, rflags = po.ref_gpu_transform(omeshes[0], ctx.obstacle_aabb(), confs)
    n, nv = len(confs), ctx.n_rob_v
    d_confs = torch.from_numpy(confs).cuda()
    dx, dy, dz = (torch.zeros(n * nv, dtype=torch.float32, device="cuda") for _ in range(3))
    d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ctx.broad_fused(d_confs, n, dx, dy, dz, None, d_valid, torch.cuda.current_stream())
    torch.cuda.synchronize()
    for got, want in ((dx, rx), (dy, ry), (dz, rz)):
        assert np.array_equal(got.cpu().numpy().reshape(n, nv).view(np.uint32), want.view(np.uint32))
    assert np.array_equal(d_valid.cpu().numpy(), rflags)
    # and the CPU oracle's libdevice sinf/cosf restatement agrees with the GPU bit for bit
    ox, oy, oz = po.transform(confs, omeshes[0].x, omeshes[0].y, omeshes[0].z)
    assert np.array_equal(ox.view(np.uint32), rx.view(np.uint32))
    assert np.array_equal(oy.view(np.uint32), ry.view(np.uint32))
    assert np.array_equal(oz.view(np.uint32), rz.view(np.uint32))


# ---- obstacle tables in global memory (meshes that do not fit in shared memory) ----------------
def test_global_obstacle_mode_same_flags(rc, meshes, mixed_confs, mixed_oracle, monkeypatch):
    """The same alpha1.0 check with the obstacle BVH/triangles read from global memory."""
    monkeypatch.setenv("RC_FORCE_GLOBAL_OBSTACLE", "1")
    c = rc.Context(meshes[0], meshes[1], device=0)
    try:
        got, bits = run_check(c, mixed_confs, rc.RC_MODE_CULLED, bitmask=True)
    finally:
        c.close()
    assert np.array_equal(got, mixed_oracle) and np.array_equal(bits, mixed_oracle)


@pytest.mark.parametrize("levels,n", [(1, 400), (2, 160)])
def test_subdivided_obstacle_matches_oracle(rc, po, meshes, levels, n):
    """BASELINE config 5 in small: robot vs the midpoint-subdivided obstacle (4,032 / 16,128
    triangles).  The reference cannot run this (its meshes live in 64 KB of __constant__ memory),
    so the checker is the CPU oracle, which takes any mesh size."""
    rob, obs = meshes[0], meshes[1].subdivide(levels)
    assert obs.n_t == 1008 * 4 ** levels
    confs = rc.createAlphaBotConfigurations(n, True, 700 + levels)
    want = po.check(po.OracleMesh(*rob.arrays()), po.OracleMesh(*obs.arrays()), confs, po.VARIANT_GPU)
    c = rc.Context(rob, obs, device=0)
    try:
        got = run_check(c, confs, rc.RC_MODE_CULLED)
        c.stats_enable(True)
        c.stats(reset=True)
        run_check(c, confs, rc.RC_MODE_CULLED)
        assert c.stats()["overflow"] == 0
    finally:
        c.close()
    assert np.array_equal(got, want)
    assert 0 < int(want.sum()) < n


@pytest.mark.parametrize("n_v,n", [(6, 2), (7, 37), (64, 33), (100, 5), (792, 9)])
def test_aabb_and_transform_stages_odd_sizes(rc, ctx, po, n_v, n):
    """generate-AABB/test-generate-AABB.cu:83-114 style: exact equality against a serial min/max,
    on sizes that exercise the scalar (unaligned) and the 128-bit paths of the stage kernels."""
    torch = _torch()
    rng = np.random.default_rng(n_v * 1000 + n)
    verts = rng.uniform(-50, 50, size=(n, n_v, 3)).astype(np.float32)
    want = np.concatenate([verts.min(1), verts.max(1)], axis=1)
    d = torch.from_numpy(verts).cuda()
    out = torch.zeros(n * 6, dtype=torch.float32, device="cuda")
    ctx.generate_aabb(d, n_v, n, out, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().reshape(n, 6), want)
    assert np.array_equal(ctx.generate_aabb_host(verts, n_v, n), want)
    # genTransformedCopies with a caller-supplied base mesh of n_v vertices
    base = verts[0].copy()
    confs = rc.createAlphaBotConfigurations(n, True, 9)
    d_base = torch.from_numpy(base).cuda()
    d_confs = torch.from_numpy(confs).cuda()
    d_out = torch.zeros(n * n_v * 3, dtype=torch.float32, device="cuda")
    ctx.transform_aos(d_confs, n, d_out, d_base=d_base, n_v=n_v, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    ox, oy, oz = po.transform(confs, base[:, 0], base[:, 1], base[:, 2])
    got = d_out.cpu().numpy().reshape(n, n_v, 3)
    assert np.array_equal(got[..., 0], ox) and np.array_equal(got[..., 1], oy) and np.array_equal(got[..., 2], oz)


# ---- MegaKernel-signature entry, planner-loop helpers ------------------------------------------
def test_megakernel_signature_and_tri_aabbs(rc, ctx, meshes, mixed_confs, mixed_oracle):
    torch = _torch()
    n = len(mixed_confs)
    d_confs = torch.from_numpy(mixed_confs).cuda()
    d_valid = torch.ones(n, dtype=torch.uint8, device="cuda")   # pre-filled true, MegaKernel.cu:448-452
    tri_boxes = ctx.obstacle_tri_aabbs()
    d_tri = torch.from_numpy(tri_boxes).cuda()
    d_obs = torch.from_numpy(ctx.obstacle_aabb()).cuda()
    ctx.megakernel(d_confs, d_valid, n, d_obs, d_tri, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(d_valid.cpu().numpy(), mixed_oracle)
    # generateTriAABBs: min/max of the three vertices of every obstacle triangle
    x, y, z, v1, v2, v3 = meshes[1].arrays()
    P = np.stack([x, y, z], 1)
    T = np.stack([P[v1], P[v2], P[v3]], 1)
    assert np.array_equal(tri_boxes, np.concatenate([T.min(1), T.max(1)], 1))


@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 100003])
def test_compaction(rc, ctx, n):
    torch = _torch()
    rng = np.random.default_rng(n)
    flags = (rng.random(n) < 0.37).astype(np.uint8)
    from robocheck_b200 import sharding as sh
    words = sh.pack_bitmask(flags, (n + 31) // 32 + 1)
    d_flags = torch.from_numpy(flags).cuda()
    d_words = torch.from_numpy(words.view(np.int32)).cuda()
    for want in (1, 0):
        truth = np.nonzero(flags == want)[0].astype(np.uint32)
        for use_mask in (False, True):
            d_idx = torch.full((max(n, 1),), -1, dtype=torch.int32, device="cuda")
            d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            ctx.compact(n, d_idx, d_cnt, d_valid=None if use_mask else d_flags,
                        d_bitmask=d_words if use_mask else None, want=want,
                        stream=torch.cuda.current_stream())
            torch.cuda.synchronize()
            cnt = int(d_cnt.item())
            assert cnt == len(truth)
            assert np.array_equal(d_idx.cpu().numpy()[:cnt].view(np.uint32), truth)


def test_device_generator_and_planner_loop(rc, ctx, po, omeshes):
    """generate on the device -> check -> compact the valid ones; flags of the generated
    configurations are checked against the CPU oracle."""
    torch = _torch()
    n = 3000
    rng_box = [-0.3, 0.3, -0.2, 0.2, -0.1, 0.1]
    d_confs = torch.zeros(n * 6, dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    ctx.generate_device(d_confs, n, rng_box, 1234, s)
    d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_mask = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    ctx.check(d_confs, n, d_valid, d_mask, rc.RC_MODE_CULLED, s)
    d_idx = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.compact(n, d_idx, d_cnt, d_bitmask=d_mask, stream=s)
    torch.cuda.synchronize()
    confs = d_confs.cpu().numpy().reshape(n, 6)
    assert np.all(confs[:, 0] >= -0.3) and np.all(confs[:, 0] < 0.3) and np.all(np.abs(confs[:, 3:]) <= np.pi)
    assert confs[:, 0].std() > 0.1 and abs(confs[:, 4].mean()) < 0.2          # looks uniform
    d2 = torch.zeros(n * 6, dtype=torch.float32, device="cuda")
    ctx.generate_device(d2, n, rng_box, 1234, s); torch.cuda.synchronize()
    assert torch.equal(d2, d_confs)                                            # reproducible
    want = po.check(omeshes[0], omeshes[1], confs, po.VARIANT_GPU)
    assert np.array_equal(d_valid.cpu().numpy(), want)
    cnt = int(d_cnt.item())
    assert np.array_equal(d_idx.cpu().numpy()[:cnt], np.nonzero(want)[0])


def test_cpp_harness_end_to_end(rc, golden, tmp_path):
    """robocheck_cli = the reference's full-integration-test main() written against the reference's
    own signatures (include/robocheck_compat.hpp): must reproduce every tag of the shipped sets."""
    import subprocess
    cli = os.path.join(rc.REPO_ROOT, "robo-check_b200", "robocheck_cli")
    for key, n_valid in (("easy_confs10k", 9997), ("hard_confs10k", 7328)):
        g = golden[key]
        conf = str(tmp_path / (key + ".conf"))
        rc.writeConfigurationToFileTagged(g["confs"], g["tags"], conf)
        back, tags = rc.readConfigurationFromFileTagged(conf)
        assert np.array_equal(back, g["confs"]) and np.array_equal(tags, g["tags"])  # 6-digit text is exact here
        out = subprocess.run([cli, conf, os.path.join(rc.ASSET_DIR, "robot.g"),
                              os.path.join(rc.ASSET_DIR, "obstacle.g")], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert f"{n_valid} valid" in out.stdout and "0 disagreements" in out.stdout
        # the same program through rc_check_multi of the C ABI (three contexts, device index modulo
        # the devices present)
        out = subprocess.run([cli, conf, os.path.join(rc.ASSET_DIR, "robot.g"),
                              os.path.join(rc.ASSET_DIR, "obstacle.g"), "--gpus", "3"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "rc_check_multi over 3 contexts" in out.stdout
        assert f"{n_valid} valid" in out.stdout and "0 disagreements" in out.stdout


# ---- sphere-approximated robot (BASELINE config 3; this repo's own definition) -----------------
def test_sphere_robot_matches_oracle(rc, po, omeshes, meshes):
    torch = _torch()
    c = rc.Context(meshes[0], meshes[1], device=0)
    try:
        confs = np.concatenate([rc.createAlphaBotConfigurations(1500, True, 801),
                                rc.createAlphaBotConfigurations(500, False, 802)]).astype(np.float32)
        n = len(confs)
        d_confs = torch.from_numpy(confs).cuda()
        d_valid = torch.full((n,), 9, dtype=torch.uint8, device="cuda")
        d_mask = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
        c.check_spheres(d_confs, n, d_valid, d_mask, torch.cuda.current_stream())
        torch.cuda.synchronize()
        spheres = c.robot_spheres()
        assert spheres.shape == (1008, 4)          # default: one enclosing sphere per robot triangle
        # every sphere really encloses its triangle
        x, y, z, v1, v2, v3 = meshes[0].arrays()
        P = np.stack([x, y, z], 1).astype(np.float64)
        for tri in (v1, v2, v3):
            assert np.all(np.linalg.norm(P[tri] - spheres[:, :3], axis=1) <= spheres[:, 3] * (1 + 1e-6))
        want = po.check_spheres(spheres, omeshes[1], confs)
        got = d_valid.cpu().numpy()
        assert np.array_equal(got, want)
        bits = np.unpackbits(d_mask.cpu().numpy().view(np.uint8), bitorder="little")[:n]
        assert np.array_equal(bits, want)
        # the spheres cover the robot surface: a configuration the spheres clear is also clear for
        # the mesh (never the other way round)
        mesh_flags = run_check(c, confs, rc.RC_MODE_CULLED)
        assert np.all(got <= mesh_flags) and 0 < int(got.sum()) < int(mesh_flags.sum())
        # a caller-supplied set: two spheres
        c.set_robot_spheres(np.array([[10, 0, 5, 2.0], [40, 10, 8, 0.5]], np.float32))
        c.check_spheres(d_confs, n, d_valid, None, torch.cuda.current_stream())
        torch.cuda.synchronize()
        want2 = po.check_spheres(c.robot_spheres(), omeshes[1], confs)
        assert np.array_equal(d_valid.cpu().numpy(), want2) and int(want2.sum()) > int(want.sum())
    finally:
        c.close()


def test_stack_overflow_fallback(rc, meshes, mixed_confs, mixed_oracle, monkeypatch):
    """The pair stack never overflows in practice; shrink it so that it does and check that the
    exhaustive fallback takes over with the same flags."""
    monkeypatch.setenv("RC_STACK_CAP", "64")
    c = rc.Context(meshes[0], meshes[1], device=0)
    try:
        c.stats_enable(True)
        c.stats(reset=True)
        got = run_check(c, mixed_confs[:600], rc.RC_MODE_CULLED)
        st = c.stats()
    finally:
        c.close()
    assert st["overflow"] > 0
    assert np.array_equal(got, mixed_oracle[:600])


def test_million_configurations_bitmask_and_sample(rc, ctx, po, omeshes):
    """Maximum-size style check (BASELINE config 4: 1M configurations of the reference's "easy"
    sampler, ragged count): bytes and bitmask agree everywhere, the counts add up, and a random
    sample of the flags -- all the interesting ones plus 1,500 others -- matches the CPU oracle."""
    n = 1_000_003
    confs = rc.createAlphaBotConfigurations(n, False, 510)
    got, bits = run_check(ctx, confs, rc.RC_MODE_CULLED, bitmask=True)
    assert np.array_equal(got, bits)
    assert set(np.unique(got)) <= {0, 1}
    rng = np.random.default_rng(1)
    sample = np.unique(np.concatenate([np.nonzero(got == 0)[0][:500], rng.choice(n, 1500, replace=False),
                                       [0, n - 1]]))
    want = po.check(omeshes[0], omeshes[1], confs[sample], po.VARIANT_GPU)
    assert np.array_equal(got[sample], want)
    assert 0 < int((got == 0).sum()) < n // 100
    # the host call at this size is the single-launch form (the kernel classifies, 31,251 batches drawn
    # by the classifying warps, lists consumed while they grow): same flags, ragged tail included
    assert np.array_equal(ctx.check_host(confs), got)
    hard = rc.createAlphaBotConfigurations(400_001, True, 512)
    assert np.array_equal(ctx.check_host(hard), run_check(ctx, hard, rc.RC_MODE_CULLED))
    assert ctx.stats()["watchdog"] == 0


@pytest.mark.parametrize("shape", ["tetra", "single", "strip7"])
def test_tiny_meshes_all_paths(rc, po, shape):
    """Degenerate sizes: one-leaf trees, fewer vertices than lanes, odd triangle counts."""
    torch = _torch()
    if shape == "tetra":
        v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32) * 2
        t = np.array([[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]], np.int32)
    elif shape == "single":
        v = np.array([[0, 0, 0], [2, 0, 0.3], [0, 2, -0.2]], np.float32)
        t = np.array([[0, 1, 2]], np.int32)
    else:
        v = np.array([[i * 0.7, (i % 2) * 1.1, 0.1 * i * i] for i in range(9)], np.float32)
        t = np.array([[i, i + 1, i + 2] for i in range(7)], np.int32)
    rob = rc.Mesh.from_arrays(v[:, 0], v[:, 1], v[:, 2], t[:, 0], t[:, 1], t[:, 2])
    ov = v * np.float32(1.3) + np.float32([0.4, -0.2, 0.3])
    obs = rc.Mesh.from_arrays(ov[:, 0], ov[:, 1], ov[:, 2], t[:, 0], t[:, 1], t[:, 2])
    confs = rc.generateConfs(700, -2.5, 2.5, -2.5, 2.5, -2.5, 2.5, 77)
    want = po.check(po.OracleMesh(*rob.arrays()), po.OracleMesh(*obs.arrays()), confs, po.VARIANT_GPU)
    assert 0 < int(want.sum()) < len(want)
    c = rc.Context(rob, obs, device=0)
    try:
        for mode in (rc.RC_MODE_CULLED, rc.RC_MODE_BRUTE):
            got, bits = run_check(c, confs, mode, bitmask=True)
            assert np.array_equal(got, want) and np.array_equal(bits, want)
        n, nv = len(confs), rob.n_v
        d_confs = torch.from_numpy(confs).cuda()
        dx, dy, dz = (torch.zeros(n * nv, dtype=torch.float32, device="cuda") for _ in range(3))
        d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
        s = torch.cuda.current_stream()
        c.broad_fused(d_confs, n, dx, dy, dz, None, d_valid, s)
        c.narrow_phase(dx, dy, dz, n, d_valid, s)
        torch.cuda.synchronize()
        assert np.array_equal(d_valid.cpu().numpy(), want)
        c.check_spheres(d_confs, n, d_valid, None, s)
        torch.cuda.synchronize()
        assert np.array_equal(d_valid.cpu().numpy(), po.check_spheres(c.robot_spheres(), po.OracleMesh(*obs.arrays()), confs))
    finally:
        c.close()


def _only_disjoint_pairs_hit(po, rob_a, obs_a, conf, variant, clearance=2e-3):
    """True iff every pair the reference arithmetic reports for `conf` has bounding boxes more than
    `clearance` apart (the triangles are disjoint; the report is a rounding artefact)."""
    wx, wy, wz = po.transform(conf[None, :], *rob_a[:3])
    w = np.stack([wx[0], wy[0], wz[0]], 1)
    ov = np.stack(obs_a[:3], 1)
    hits = 0
    for a, b, c3 in zip(*rob_a[3:]):
        r = w[[a, b, c3]]
        for e, f, g in zip(*obs_a[3:]):
            o = ov[[e, f, g]]
            if po.pair_hits(r.reshape(-1), o.reshape(-1), variant)[0]:
                hits += 1
                if np.maximum(r.min(0) - o.max(0), o.min(0) - r.max(0)).max() <= clearance:
                    return False
    return hits > 0


def test_coplanar_mode_all_paths(rc, po):
    """RC_COPLANAR_SKIP reproduces the reference (undecided pairs never collide,
    narrow-phase.cu:647-652,:680), RC_COPLANAR_DECIDE decides them in the common plane; both modes
    bit-equal to the oracle in the exhaustive and stage-wise paths, on a scene where the option
    changes more than 100 flags (tests/coplanar_scene.py).

    The scene is made of large triangles lying in one plane, which is also where the reference's
    arithmetic breaks down: for nearly parallel planes its intersection line is rounding noise and it
    reports some pairs whose triangles are far apart inside the plane (configuration 16: boxes 0.39
    apart).  The tree-culled kernel never looks at such a pair (culling safety), so for
    RC_MODE_CULLED the assertion is: any flag that differs is valid-here / invalid-there, and every
    pair the oracle reports for that configuration has boxes more than the culling clearance apart."""
    torch = _torch()
    from tests import coplanar_scene
    rob_a, obs_a, confs = coplanar_scene.build(po)
    R, O = po.OracleMesh(*rob_a), po.OracleMesh(*obs_a)
    variant = {rc.RC_COPLANAR_SKIP: po.VARIANT_GPU, rc.RC_COPLANAR_DECIDE: po.VARIANT_GPU | po.FLAG_COPLANAR}
    want = {m: po.check(R, O, confs, v) for m, v in variant.items()}
    assert int((want[0] != want[1]).sum()) > 100
    c = rc.Context(rc.Mesh.from_arrays(*rob_a), rc.Mesh.from_arrays(*obs_a), device=0)
    try:
        assert c.coplanar_mode() == rc.RC_COPLANAR_SKIP  # the default is the reference's behaviour
        with pytest.raises(rc.RoboCheckError):
            c.set_coplanar_mode(7)
        n, nv = len(confs), len(rob_a[0])
        d_confs = torch.from_numpy(confs).cuda()
        s = torch.cuda.current_stream()
        for mode in (rc.RC_COPLANAR_DECIDE, rc.RC_COPLANAR_SKIP):
            c.set_coplanar_mode(mode)
            assert c.coplanar_mode() == mode
            got, bits = run_check(c, confs, rc.RC_MODE_BRUTE, bitmask=True)
            assert np.array_equal(got, want[mode]), (mode, int((got != want[mode]).sum()))
            assert np.array_equal(bits, want[mode])
            dx, dy, dz = (torch.zeros(n * nv, dtype=torch.float32, device="cuda") for _ in range(3))
            d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
            c.broad_fused(d_confs, n, dx, dy, dz, None, d_valid, s)
            c.narrow_phase(dx, dy, dz, n, d_valid, s)
            torch.cuda.synchronize()
            assert np.array_equal(d_valid.cpu().numpy(), want[mode])
            # tree-culled kernel, with and without statistics, and through the host entry point
            culled, bits = run_check(c, confs, rc.RC_MODE_CULLED, bitmask=True)
            assert np.array_equal(bits, culled)
            differ = np.nonzero(culled != want[mode])[0]
            assert len(differ) <= 3 and np.all(culled[differ] == 1)
            for i in differ:
                assert _only_disjoint_pairs_hit(po, rob_a, obs_a, confs[i], variant[mode])
            # the coplanar decisions themselves all survive the culling
            changed = want[0] != want[1]
            assert np.array_equal(culled[changed], want[mode][changed])
            c.stats_enable(True)
            again = run_check(c, confs, rc.RC_MODE_CULLED)
            c.stats_enable(False)
            assert np.array_equal(again, culled) and c.stats(reset=True)["watchdog"] == 0
            assert np.array_equal(np.asarray(c.check_host(confs)), culled)
    finally:
        c.close()


@pytest.mark.parametrize("workload", ["C2_hard_100k", "C4_uniform_1M", "hard_1M"])
def test_full_size_workloads_every_flag(rc, ctx, po, omeshes, workload):
    """BASELINE.json's sizes, every flag checked: the tree-culled CPU checker (oracle_bvh.c, pinned
    against the exhaustive oracle in tests/test_oracle.py) decides all of the bench workload
    (configs[1], seed 508), all of configs[3] (1M of the reference's sampler) and 1M hard ones."""
    n, hard, seed = {"C2_hard_100k": (100_000, True, 508), "C4_uniform_1M": (1_000_000, False, 510),
                     "hard_1M": (1_000_003, True, 511)}[workload]
    confs = rc.createAlphaBotConfigurations(n, hard, seed)
    want, st = po.check_bvh(omeshes[0], omeshes[1], confs, po.VARIANT_GPU, want_stats=True)
    got, bits = run_check(ctx, confs, rc.RC_MODE_CULLED, bitmask=True)
    assert np.array_equal(got, want), int((got != want).sum())
    assert np.array_equal(bits, want)
    assert st["narrow_invalid"] > 0 and st["broad_cleared"] > 0
    host = ctx.check_host(confs)
    assert np.array_equal(host, want)


@pytest.mark.parametrize("levels,n", [(2, 200_000), (5, 100_000)])
def test_subdivided_obstacle_full_size_every_flag(rc, po, meshes, levels, n):
    """BASELINE configs[4] at its real obstacle sizes (16,128 and 1,032,192 triangles, tables in
    global memory), "easy" configurations with the 10 % "hard" admixture of SURVEY 8(d): every flag
    against the tree-culled CPU checker."""
    rob, obs = meshes[0], meshes[1].subdivide(levels)
    assert obs.n_t == 1008 * 4 ** levels
    confs = np.concatenate([rc.createAlphaBotConfigurations(n * 9 // 10, False, 511),
                            rc.createAlphaBotConfigurations(n // 10, True, 512)]).astype(np.float32)
    want, st = po.check_bvh(po.OracleMesh(*rob.arrays()), po.OracleMesh(*obs.arrays()), confs,
                            po.VARIANT_GPU, want_stats=True)
    assert st["narrow_invalid"] > 1000 and st["narrow_valid"] > 1000
    c = rc.Context(rob, obs, device=0)
    try:
        got, bits = run_check(c, confs, rc.RC_MODE_CULLED, bitmask=True)
        assert np.array_equal(got, want), int((got != want).sum())
        assert np.array_equal(bits, want)
    finally:
        c.close()


def test_pipelined_host_calls(rc, ctx, po, omeshes):
    """rc_check_host_submit / rc_check_host_wait: batches of different sizes in flight two at a time,
    page-locked and pageable buffers, both kernels, waits in and out of order -- the flags equal the
    synchronous call's (which equal the oracle's on the first batch)."""
    torch = _torch()
    sizes = [5000, 40001, 33, 70000, 1, 9999, 20000]
    batches = [rc.createAlphaBotConfigurations(n, k % 3 != 2, 900 + k) for k, n in enumerate(sizes)]
    want = [ctx.check_host(b) for b in batches]
    assert np.array_equal(want[0], po.check_bvh(omeshes[0], omeshes[1], batches[0], po.VARIANT_GPU))
    for pinned in (True, False):
        if pinned:
            ins = [torch.from_numpy(b).pin_memory().numpy() for b in batches]
            outs = [torch.full((len(b),), 9, dtype=torch.uint8).pin_memory().numpy() for b in batches]
        else:
            ins = [b.copy() for b in batches]
            outs = [np.full(len(b), 9, np.uint8) for b in batches]
        tickets = []
        for k in range(len(batches)):
            mode = rc.RC_MODE_BRUTE if k == 2 else rc.RC_MODE_CULLED
            tickets.append(ctx.check_host_submit(ins[k], outs[k], mode))
            if k >= 1 and k % 2 == 1:
                ctx.check_host_wait(tickets[k - 1])       # one behind: two batches in flight
        for t in reversed(tickets):                        # the rest, newest first; finished ones are no-ops
            ctx.check_host_wait(t)
        assert tickets == sorted(tickets) and len(set(tickets)) == len(tickets)
        for k in range(len(batches)):
            assert np.array_equal(outs[k], want[k]), (pinned, k, int((outs[k] != want[k]).sum()))
    with pytest.raises(rc.RoboCheckError):
        ctx.check_host_wait(10 ** 6)
    assert ctx.stats()["watchdog"] == 0


# ---- round 2: the exact BASELINE workloads against exhaustive ground truth ---------------------
def test_bench_workload_vs_reference_gpu_pipeline_and_brute(rc, ctx, po):
    """The bench workload itself (BASELINE configs[1]: 100,000 hard configurations, seed 508) against
    the UNMODIFIED reference pipeline run on this GPU (broadPhaseFused_sep from oracle/_ref, an
    exhaustive sweep with the reference's own kernels) and against RC_MODE_BRUTE: no checker in this
    test shares the culling of the kernel under test."""
    confs = rc.createAlphaBotConfigurations(100_000, True, 508)
    culled, bits = run_check(ctx, confs, rc.RC_MODE_CULLED, bitmask=True)
    brute = run_check(ctx, confs, rc.RC_MODE_BRUTE)
    assert int((culled != brute).sum()) == 0 and np.array_equal(bits, culled)
    assert np.array_equal(ctx.check_host(confs), brute)
    if not po.ref_available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    from tests import refpipe
    want = refpipe.oracle_gpu_flags(rc, po, confs)
    assert int((culled != want).sum()) == 0
    assert int((brute != want).sum()) == 0


def test_c4_first_300k_vs_reference_gpu_pipeline(rc, ctx, po):
    """BASELINE configs[3] (1 M of the reference's sampler, seed 510): the first 300,000 against the
    unmodified reference pipeline on this GPU, all 1 M culled == brute."""
    confs = rc.createAlphaBotConfigurations(1_000_000, False, 510)
    culled = run_check(ctx, confs, rc.RC_MODE_CULLED)
    brute = run_check(ctx, confs, rc.RC_MODE_BRUTE)
    assert int((culled != brute).sum()) == 0
    if not po.ref_available():
        pytest.skip("oracle/_ref/libref.so not built")
    from tests import refpipe
    want = refpipe.oracle_gpu_flags(rc, po, confs[:300_000])
    assert int((culled[:300_000] != want).sum()) == 0


def test_c5_level5_exhaustive_spot_check(rc, po, meshes):
    """BASELINE configs[4] at its real obstacle size (1,032,192 triangles): 20 configurations -- near
    contact, where the narrow phase decides -- against the EXHAUSTIVE CPU sweep (po.check: all
    1008 x 1,032,192 pairs, no tree anywhere), plus their hitting-pair counts against the tree
    descent's."""
    torch = _torch()
    rob, obs = meshes[0], meshes[1].subdivide(5)
    assert obs.n_t == 1008 * 4 ** 5
    confs = np.concatenate([rc.createAlphaBotConfigurations(14, True, 512),
                            rc.createAlphaBotConfigurations(6, False, 511)]).astype(np.float32)
    R, O = po.OracleMesh(*rob.arrays()), po.OracleMesh(*obs.arrays())
    want, want_counts = po.check_counts(R, O, confs)
    assert 0 < int(want.sum()) < len(want)
    c = rc.Context(rob, obs, device=0)
    try:
        got = run_check(c, confs, rc.RC_MODE_CULLED)
        assert np.array_equal(got, want)
        d_confs = torch.from_numpy(confs).cuda()
        d_cnt = torch.zeros(len(confs), dtype=torch.int32, device="cuda")
        c.hit_counts(d_confs, len(confs), d_cnt, rc.RC_MODE_CULLED, stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert np.array_equal(d_cnt.cpu().numpy().view(np.uint32), want_counts)
    finally:
        c.close()


def test_hit_pair_counts(rc, ctx, po, omeshes):
    """SURVEY appendix A, closing rule: per-configuration number of hitting pairs from a sweep without
    early exit.  Equal flags could hide compensating per-pair errors (a colliding configuration has
    ~100 hitting pairs); equal COUNTS cannot.  (a) exhaustive kernel == exhaustive CPU oracle on 600
    configurations, exactly; (b) tree descent vs exhaustive kernel on the whole bench workload: the
    tree never reports a pair the sweep does not, and the pairs it skips are the reference
    arithmetic's artefacts on triangles that lie apart (profiles/tools/count_diff.py lists them: 5 of
    2,926,009 reported pairs on this set, each in a configuration with > 100 other hits) -- a handful
    per million, never a flag; (c) rc_audit reports the same numbers."""
    torch = _torch()
    s = torch.cuda.current_stream()
    confs = rc.createAlphaBotConfigurations(100_000, True, 508)
    n = len(confs)
    d_confs = torch.from_numpy(confs).cuda()
    d_cb = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    d_cc = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    d_fb = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    d_fc = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    ctx.hit_counts(d_confs, n, d_cb, rc.RC_MODE_BRUTE, d_fb, s)
    ctx.hit_counts(d_confs, n, d_cc, rc.RC_MODE_CULLED, d_fc, s)
    torch.cuda.synchronize()
    cb, cc = d_cb.cpu().numpy().view(np.uint32), d_cc.cpu().numpy().view(np.uint32)
    fb, fc = d_fb.cpu().numpy(), d_fc.cpu().numpy()
    m = 600
    want_flags, want_counts = po.check_counts(omeshes[0], omeshes[1], confs[:m])
    assert np.array_equal(fb[:m], want_flags)
    assert np.array_equal(cb[:m], want_counts), int((cb[:m] != want_counts).sum())
    assert int((fc != fb).sum()) == 0
    assert not np.any(cc == 0xffffffff)                   # no descent fell back to the exhaustive sweep
    gap = cb.astype(np.int64) - cc.astype(np.int64)
    assert gap.min() >= 0                                  # the tree reports a subset of the sweep's pairs
    assert gap.sum() <= 1e-5 * cb.sum() and int((gap != 0).sum()) <= 20, (int(gap.sum()), int((gap != 0).sum()))
    assert np.all(cc[gap != 0] > 50)                       # and only where plenty of other pairs hit
    assert np.array_equal(cb > 0, fb == 0)
    assert 50 < cb[fb == 0].mean() < 300      # ~109 hitting pairs per colliding configuration
    # the flags of the counting kernels are the flags of the production kernels
    assert np.array_equal(fb, run_check(ctx, confs, rc.RC_MODE_CULLED))
    nf, nc, idx = ctx.audit(d_confs, n)
    assert nf == 0 and nc == int((gap != 0).sum()) and np.array_equal(idx, np.nonzero(gap)[0][:len(idx)])
    assert ctx.stats()["overflow"] == 0


def test_audit_reports_the_coplanar_scene(rc, po):
    """rc_audit on the scene where the exactness contract of RC_MODE_CULLED does NOT hold
    (tests/coplanar_scene.py: the reference arithmetic reports triangle pairs that lie far apart in a
    common plane): the audit must say so -- counts differ -- instead of 0."""
    torch = _torch()
    from tests import coplanar_scene
    rob_a, obs_a, confs = coplanar_scene.build(po)
    c = rc.Context(rc.Mesh.from_arrays(*rob_a), rc.Mesh.from_arrays(*obs_a), device=0)
    try:
        d_confs = torch.from_numpy(confs).cuda()
        nf, nc, idx = c.audit(d_confs, len(confs), cap=16)
        culled = run_check(c, confs, rc.RC_MODE_CULLED)
        brute = run_check(c, confs, rc.RC_MODE_BRUTE)
        assert nf == int((culled != brute).sum())
        assert nc >= nf and nc > 0 and len(idx) > 0
    finally:
        c.close()


def test_device_known_answer_tests(rc, ctx, po, omeshes):
    """The reference's helper KATs (narrow-phase/test-narrow-phase.cu:7-626) run on the DEVICE
    functions the kernels are built from (plane_of, sdist, straddles, is_coplanar, la_solve,
    tri_interval ...), and every intermediate of several thousand real alpha1.0 pairs compared bit for bit with
    the CPU oracle's helper chain."""
    PTS = np.array([[1, 2, 3], [2, 2, 3.5], [2, 2.5, 3], [1, 3, 4], [4, 7, 2], [5, 1, 3],
                    [100, 300, 400], [400, 700, 200], [500, 100, 300]], np.float32)
    T1, T2, T3 = PTS[0:3].reshape(-1), PTS[3:6].reshape(-1), PTS[6:9].reshape(-1)
    out, verdict = ctx.kat_pairs(np.stack([T1, T1, T1]), np.stack([T2, T3, T1]))
    # test-narrow-phase.cu:85-160: plane of T1 = (-0.25, 0.5, 0.5 | -2.25), distances of T2 = 1, 1.25, -1.5
    assert out[0, 0:4].tolist() == [-0.25, 0.5, 0.5, -2.25]
    assert out[0, 4:7].tolist() == [1.0, 1.25, -1.5]
    # :162-214 no_overlap: T2 straddles the plane of T1, T3 and T1 itself do not
    assert verdict[0] & 1 == 1 and verdict[1] & 1 == 0 and verdict[2] & 1 == 0
    # :522-575 parametric values of the pair (T1, T2) on the intersection line (the unit test does not
    # ask whether T1 straddles the plane of T2; the device test harness evaluates the tail regardless)
    want_t = [108.607147, 143.744751, 130.328568, 88.860390]
    assert np.allclose(out[0, 20:24], want_t, rtol=1e-6)
    assert (verdict[0] >> 2) == 1                      # the intervals overlap
    # :216-246 la_solve
    ls = ctx.kat_la_solve([[2, 3, 4, 5, 12, 22], [0, 3, 4, 5, 12, 24]])
    assert np.allclose(ls, [[3, 2], [1, 4]], rtol=1e-6)
    assert ls[0].tolist() == list(po.la_solve_gpu(2, 3, 4, 5, 12, 22))
    # :577-626 coplanarity: a triangle against a copy of itself shifted inside its own plane
    shifted = (PTS[0:3] + (PTS[1] - PTS[0]) * np.float32(0.25)).reshape(-1).astype(np.float32)
    _, vc = ctx.kat_pairs(T1[None], shifted[None])
    assert (vc[0] >> 2) == 2 and po.pair_intermediates(T1, shifted)[1] >> 2 == 2
    # real pairs: every robot triangle of four near-contact configurations against a stride of the obstacle
    rob, obs = omeshes
    confs = rc.createAlphaBotConfigurations(4, True, 508)
    wx, wy, wz = po.transform(confs, rob.x, rob.y, rob.z)
    ov = np.stack([obs.x, obs.y, obs.z], 1)
    o9_all = np.concatenate([ov[obs.v1], ov[obs.v2], ov[obs.v3]], 1)          # [n_obs_t, 9]
    r9s, o9s = [], []
    rng = np.random.default_rng(3)
    for ci in range(len(confs)):
        w = np.stack([wx[ci], wy[ci], wz[ci]], 1)
        r9_all = np.concatenate([w[rob.v1], w[rob.v2], w[rob.v3]], 1)
        # nearby pairs are the interesting ones: pick, for a sample of robot triangles, the obstacle
        # triangles whose boxes overlap, plus random ones
        for j in rng.choice(len(r9_all), 400, replace=False):
            r = r9_all[j].reshape(3, 3)
            lo, hi = r.min(0) - 0.05, r.max(0) + 0.05
            o = o9_all.reshape(-1, 3, 3)
            near = np.nonzero(np.all(o.min(1) <= hi, 1) & np.all(o.max(1) >= lo, 1))[0][:32]
            for k in list(near) + list(rng.choice(len(o9_all), 4, replace=False)):
                r9s.append(r9_all[j]); o9s.append(o9_all[k])
    r9s, o9s = np.array(r9s, np.float32), np.array(o9s, np.float32)
    assert len(r9s) > 4_000
    out, verdict = ctx.kat_pairs(r9s, o9s)
    n_tail = n_hit = 0
    for i in range(len(r9s)):
        w_out, w_v = po.pair_intermediates(r9s[i], o9s[i])
        both = (w_v & 3) == 3
        # planes, signed distances and the two plane tests: always; intersection line, parametric values
        # and the verdict of the tail: where the kernels evaluate them (both plane tests passed)
        assert (verdict[i] & 3) == (w_v & 3), (i, int(verdict[i]), w_v)
        assert np.array_equal(out[i, :14].view(np.uint32), w_out[:14].view(np.uint32)), (i, out[i], w_out)
        if both:
            assert verdict[i] == w_v, (i, int(verdict[i]), w_v)
            assert np.array_equal(out[i].view(np.uint32), w_out.view(np.uint32)), (i, out[i], w_out)
        n_tail += both
        n_hit += both and (w_v >> 2) == 1
    assert n_tail > 60 and n_hit > 5         # the sample reaches the intersection-line code and finds hits


@pytest.mark.parametrize("knobs", [{}, {"RC_SPLIT_MAX": "1"}, {"RC_SPLIT_MAX": "4", "RC_SPLIT_LOAD": "64"}])
def test_planner_sized_batches_split_descents(rc, meshes, knobs, monkeypatch):
    """Short work lists: a configuration is cut into parts (subsets of the seed pairs) that warps of
    the whole grid take; the flags are written optimistically (valid first, a part with a hit
    overwrites).  Every size from one configuration up, against the exhaustive kernel, repeated (the
    scheduling is timing dependent)."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    confs = rc.createAlphaBotConfigurations(12_000, True, 508)
    c = rc.Context(meshes[0], meshes[1], device=0)
    try:
        want = run_check(c, confs, rc.RC_MODE_BRUTE)
        for n in (1, 2, 31, 33, 100, 500, 2000, 4736, 10_000, 12_000):
            for rep in range(3):
                got, bits = run_check(c, confs[:n], rc.RC_MODE_CULLED, bitmask=True)
                assert np.array_equal(got, want[:n]), (n, rep, int((got != want[:n]).sum()))
                assert np.array_equal(bits, want[:n])
            assert np.array_equal(c.check_host(confs[:n]), want[:n])
        assert c.stats()["watchdog"] == 0
    finally:
        c.close()


def test_calls_on_different_streams_share_the_scratch_safely(rc, ctx):
    """One context driven from two streams back to back without a host synchronisation in between
    (rc_check on a side stream, then the host call on the context's own stream, then rc_check on the
    default stream): the library orders the device work, every result is right."""
    torch = _torch()
    confs = rc.createAlphaBotConfigurations(60_000, True, 321)
    want = run_check(ctx, confs, rc.RC_MODE_BRUTE)
    d_confs = torch.from_numpy(confs).cuda()
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(4):
        v1 = torch.full((len(confs),), 7, dtype=torch.uint8, device="cuda")
        v3 = torch.full((len(confs),), 7, dtype=torch.uint8, device="cuda")
        ctx.check(d_confs, len(confs), v1, None, rc.RC_MODE_CULLED, side)
        v2 = ctx.check_host(confs[:30_000])
        ctx.check(d_confs, len(confs), v3, None, rc.RC_MODE_CULLED, torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert np.array_equal(v1.cpu().numpy(), want)
        assert np.array_equal(v2, want[:30_000])
        assert np.array_equal(v3.cpu().numpy(), want)


def test_scaled_and_translated_scene(rc, meshes):
    """The culling margin scales with the scene (a cell measured in millimetres, far from the origin):
    alpha1.0 blown up 40 times and moved 3,000 units away -- culled == exhaustive on every flag."""
    x, y, z, v1, v2, v3 = meshes[0].arrays()
    ox, oy, oz, o1, o2, o3 = meshes[1].arrays()
    k, t = np.float32(40.0), np.float32([3000.0, -2000.0, 1000.0])
    obs = rc.Mesh.from_arrays(ox * k + t[0], oy * k + t[1], oz * k + t[2], o1, o2, o3)
    rob = rc.Mesh.from_arrays(x * k, y * k, z * k, v1, v2, v3)
    base = rc.createAlphaBotConfigurations(6000, True, 99)
    confs = base.copy()
    confs[:, :3] = base[:, :3] * k + t          # same relative poses, in the big frame
    c = rc.Context(rob, obs, device=0)
    try:
        a = run_check(c, confs, rc.RC_MODE_CULLED)
        b = run_check(c, confs, rc.RC_MODE_BRUTE)
    finally:
        c.close()
    assert int((a != b).sum()) == 0
    assert 0.1 < b.mean() < 0.95


def test_gather_two_ranks_on_one_device(rc, ctx, meshes):
    """rc_check_gather's device path on ONE GPU (the multi-GPU tests are skipped on a 1-GPU box): two
    contexts play rank 0 and rank 1, their gather buffers are two plain device allocations that are
    each other's "peer".  Kernels that wait for one another must not share a GPU, so the two-phase
    form is used: begin for rank 0, begin for rank 1 (each pushes its slot into both buffers and
    signals), then end for both (the waits find the signals already there).  Three calls (both gather
    areas, call numbers), ragged sizes with a short last shard, then a FRESH context on the used
    buffers (the numbering continues from the buffer's header), then the one-call form with world 1."""
    torch = _torch()
    from robocheck_b200 import sharding as sh
    s = torch.cuda.current_stream()
    n_all = 50_001
    confs = np.concatenate([rc.createAlphaBotConfigurations(n_all // 2, True, 41),
                            rc.createAlphaBotConfigurations(n_all - n_all // 2, False, 42)])
    confs = confs[np.random.default_rng(8).permutation(n_all)]
    single = run_check(ctx, confs, rc.RC_MODE_CULLED)
    world = 2
    words = sh.words_per_shard(n_all, world)
    total = int(rc.lib().rc_gather_buffer_words(world, words))
    bufs = [torch.zeros(total, dtype=torch.int32, device="cuda") for _ in range(world)]
    ptrs = [b.data_ptr() for b in bufs]
    ranks = [rc.Context(meshes[0], meshes[1], device=0) for _ in range(world)]

    def one_call(ctxs, m):
        offs = []
        for r in range(world):
            lo, hi = min(m, r * words * 32), min(m, (r + 1) * words * 32)
            d = torch.from_numpy(np.ascontiguousarray(confs[lo:hi])).cuda() if hi > lo else None
            offs.append(ctxs[r].check_gather_begin(d, hi - lo, None, ptrs, r, words, s))
        for r in range(world):
            ctxs[r].gather_end(ptrs, r, words, s)
        torch.cuda.synchronize()
        for r in range(world):
            got = sh.unpack_bitmask(bufs[r][offs[r]:offs[r] + world * words].cpu().numpy(), world * words * 32)[:m]
            assert np.array_equal(got, single[:m]), (r, m, int((got != single[:m]).sum()))

    try:
        for m in (n_all, n_all - 4097, 70):
            one_call(ranks, m)
        for r in range(world):
            assert ranks[r].gather_status(ptrs[r], world, words) == (3, 0)
            assert ranks[r].stats()["watchdog"] == 0
        fresh = [rc.Context(meshes[0], meshes[1], device=0) for _ in range(world)]
        try:
            one_call(fresh, 30_000)
            for r in range(world):
                assert fresh[r].gather_status(ptrs[r], world, words) == (4, 0)
        finally:
            for c in fresh:
                c.close()
        # world = 1: nothing to push or wait for, the result is the own slot
        w1 = (n_all + 31) // 32
        b1 = torch.zeros(int(rc.lib().rc_gather_buffer_words(1, w1)), dtype=torch.int32, device="cuda")
        d_all = torch.from_numpy(confs).cuda()
        off = ranks[0].check_gather(d_all, n_all, None, [b1.data_ptr()], 0, w1, s)
        torch.cuda.synchronize()
        assert np.array_equal(sh.unpack_bitmask(b1[off:off + w1].cpu().numpy(), w1 * 32)[:n_all], single)
    finally:
        for c in ranks:
            c.close()


@pytest.mark.parametrize("feed", ["0", "1", "2"])
def test_single_launch_form_every_size(rc, meshes, feed, monkeypatch):
    """RC_FEED: the descent kernel classifies the configurations itself (rc_culled.cuh: FeedArgs).
    0 = never (two launches, host buffers copied first), 1 = the host call reads the caller's buffer
    in place over PCIe (default), 2 = device-resident checks use the form too.  Every size from one
    configuration up -- one batch, ragged batches, split descents, a list longer than the grid -- against
    the exhaustive kernel; page-locked and pageable host buffers; repeated calls on one context (the
    list slots are tagged with the call number, not cleared)."""
    torch = _torch()
    monkeypatch.setenv("RC_FEED", feed)
    confs = np.concatenate([rc.createAlphaBotConfigurations(50_000, True, 508),
                            rc.createAlphaBotConfigurations(20_000, False, 510)])
    confs = confs[np.random.default_rng(3).permutation(len(confs))]
    c = rc.Context(meshes[0], meshes[1], device=0)
    try:
        want = run_check(c, confs, rc.RC_MODE_BRUTE)
        pinned = torch.from_numpy(confs).pin_memory()
        for n in (1, 31, 32, 33, 1000, 4737, 30_000, 70_000):
            for rep in range(2):
                got, bits = run_check(c, confs[:n], rc.RC_MODE_CULLED, bitmask=True)
                assert np.array_equal(got, want[:n]), (feed, n, rep, int((got != want[:n]).sum()))
                assert np.array_equal(bits, want[:n])
                assert np.array_equal(c.check_host(confs[:n]), want[:n]), (feed, n, rep, "pageable")
                assert np.array_equal(c.check_host(pinned.numpy()[:n]), want[:n]), (feed, n, rep, "pinned")
        # the reference's far sampler: nearly everything cleared by the classifying warps themselves
        far = rc.createAlphaBotConfigurations(80_000, False, 511)
        assert np.array_equal(c.check_host(far), run_check(c, far, rc.RC_MODE_BRUTE))
        assert c.stats()["watchdog"] == 0
    finally:
        c.close()
