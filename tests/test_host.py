"""CPU tests of the host side: the C-ABI library loads and exports every symbol
include/robocheck.h declares, mesh loaders, .conf I/O, generators, subdivision, error codes.
No compute calls: there is no GPU here and the product has no CPU fallback."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(rc):
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "robocheck.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    lib = ctypes.CDLL(rc.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == {name for name, _, _ in rc.ABI}, "python binding and header disagree"


def test_compat_symbols_link():
    """The reference's own C++ signatures (include/robocheck_compat.hpp) are exported, and the
    C++ harness written against them links (robocheck_cli is built by the Makefile)."""
    out = subprocess.run(["nm", "-DC", os.path.join(ROOT, "robo-check_b200", "librobocheck.so")],
                         capture_output=True, text=True, check=True).stdout
    for sig in ("broadPhaseFused_sep(std::vector<Configuration", "CallMegaKernel(std::vector<Configuration",
                "generateAABB(Vector3f*, unsigned int, unsigned int, AABB*)", "broadPhase(int, AABB const*",
                "narrowPhase(int, int, int, int, int, Triangle const*", "loadOBJFile(char const*",
                "readConfigurationFromFile(", "createAlphaBotConfigurations("):
        assert sig in out, sig
    assert os.path.exists(os.path.join(ROOT, "robo-check_b200", "robocheck_cli"))


def test_no_cpu_fallback(rc, meshes):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rc.RoboCheckError) as e:
        rc.Context(meshes[0], meshes[1])
    assert e.value.code == rc.RC_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may mention it."""
    pkg = os.path.join(ROOT, "robo-check_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_alpha_assets(rc, meshes):
    rob, obs = meshes
    assert rob.counts() == (792, 1008) and obs.counts() == (792, 1008)  # models/alpha1.0/README.txt:16
    # obstacle AABB quoted in SURVEY 8a (generateAABBBaseline)
    assert np.allclose(obs.aabb(), [-1.8928, -3.17052, -4.36586, 58.7349, 39.457, 10.6521], atol=1e-4)
    # robot extents hard-coded by the reference generator (Utils_rai.cu:113-118)
    assert np.allclose(rob.aabb(), [3.72119, -11.0518, -0.608012, 65.9453, 26.0984, 18.6984], atol=1e-4)
    for m in (rob, obs):
        for v in (m.v1, m.v2, m.v3):
            assert v.min() >= 0 and v.max() < m.n_v


def test_obj_loader(rc, tmp_path):
    p = tmp_path / "m.obj"
    p.write_text("# comment\no thing\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0.5\nvn 0 0 1\nvt 0 0\n"
                 "f 1//1 2//1 3//1\nf 1/1/1 2/1/1 3/1/1 4/1/1\n\nf 4 3 2\n")
    x, y, z, v1, v2, v3 = rc.loadOBJFile(str(p))
    assert x.tolist() == [0, 1, 1, 0] and z.tolist() == [0, 0, 0, 0.5]
    # `a//n` and `a/t/n` tokens, 1-based -> 0-based, quad fanned around its first corner
    assert list(zip(v1, v2, v3)) == [(0, 1, 2), (0, 1, 2), (0, 2, 3), (3, 2, 1)]
    with pytest.raises(rc.RoboCheckError) as e:
        rc.Mesh.from_obj(str(tmp_path / "missing.obj"))
    assert e.value.code == rc.RC_ERR_IO
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(rc.RoboCheckError) as e:
        rc.Mesh.from_obj(str(bad))
    assert e.value.code == rc.RC_ERR_FORMAT


def test_byu_loader_and_round_trips(rc, meshes, tmp_path):
    p = tmp_path / "t.g"
    p.write_text("1 4 2 6\n1 2\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n1 2 -3\n1 3 -4\n")
    m = rc.Mesh.from_byu(str(p))
    assert m.counts() == (4, 2)
    assert list(zip(m.v1, m.v2, m.v3)) == [(0, 1, 2), (0, 2, 3)]
    # quads are fan-split (the PMPL loader the reference carries would throw)
    q = tmp_path / "q.g"
    q.write_text("1 4 1 4\n1 1\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n1 2 3 -4\n")
    assert rc.Mesh.from_byu(str(q)).counts() == (4, 2)
    with pytest.raises(rc.RoboCheckError):
        rc.Mesh.from_byu(str(tmp_path / "nope.g"))
    trunc = tmp_path / "trunc.g"
    trunc.write_text("1 4 2 6\n1 2\n0 0 0\n1 0 0\n")
    with pytest.raises(rc.RoboCheckError) as e:
        rc.Mesh.from_byu(str(trunc))
    assert e.value.code == rc.RC_ERR_FORMAT
    # OBJ <-> BYU round trips keep every float bit
    rob = meshes[0]
    rob.save_obj(str(tmp_path / "r.obj"))
    rob.save_byu(str(tmp_path / "r.g"))
    for back in (rc.Mesh.from_obj(str(tmp_path / "r.obj")), rc.Mesh.from_byu(str(tmp_path / "r.g"))):
        for a, b in zip(rob.arrays(), back.arrays()):
            assert np.array_equal(a, b)


def test_conf_io(rc, golden, tmp_path):
    confs = golden["hard_confs10k"]["confs"][:257]
    tags = golden["hard_confs10k"]["tags"][:257]
    p = str(tmp_path / "c.conf")
    rc.writeConfigurationToFileTagged(confs, tags, p)
    head = open(p).readline()
    assert head == f"There are {int(tags.sum())} valid configurations and {int(257 - tags.sum())} invalid configurations\n"
    back, tback = rc.readConfigurationFromFileTagged(p)
    assert np.array_equal(tback, tags)
    # the reference writes 6 significant digits (default ostream precision), so the text
    # quantises; what was read from a 6-digit file survives a second round trip exactly
    assert np.allclose(back, confs, rtol=1e-5, atol=1e-6)
    rc.writeConfigurationToFileTagged(back, tback, p)
    assert np.array_equal(rc.readConfigurationFromFile(p), back)
    empty = str(tmp_path / "e.conf")
    rc.writeConfigurationToFileTagged(np.zeros((0, 6), np.float32), None, empty)
    assert rc.readConfigurationFromFile(empty).shape == (0, 6)
    with pytest.raises(rc.RoboCheckError):
        rc.readConfigurationFromFile(str(tmp_path / "missing.conf"))


def test_generators(rc):
    a = rc.createAlphaBotConfigurations(1000, True, 508)
    b = rc.createAlphaBotConfigurations(1000, True, 508)
    c = rc.createAlphaBotConfigurations(1000, True, 509)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    rng = np.array([65.9453 - 3.72119, 26.0984 + 11.0518, 18.6984 + 0.608012], np.float32)
    assert np.all(np.abs(a[:, :3]) <= rng / 200 + 1e-6)   # hard: +-range/200 (Utils_rai.cu:124-128)
    assert np.all(np.abs(a[:, 3:]) <= np.pi + 1e-6)
    e = rc.createAlphaBotConfigurations(1000, False, 1)
    assert np.all(np.abs(e[:, :3]) <= rng * 10 + 1e-3) and np.abs(e[:, 0]).max() > 100
    # draw order x, y, z, pitch, yaw, roll from ONE mt19937 stream (Utils_rai.cu:337-346):
    # a prefix of a longer run is the shorter run
    assert np.array_equal(rc.createAlphaBotConfigurations(10, True, 508), a[:10])
    g = rc.generateConfs(5, -1, 1, -2, 2, -3, 3, 7)
    assert g.shape == (5, 6) and np.all(np.abs(g[:, 1]) <= 2)


def test_subdivide(rc, meshes):
    obs = meshes[1]
    s1 = obs.subdivide(1)
    assert s1.n_t == 4 * obs.n_t
    # one new vertex per distinct edge: between 3T/2 (closed manifold, every edge shared) and 3T
    assert obs.n_v + 3 * obs.n_t // 2 <= s1.n_v <= obs.n_v + 3 * obs.n_t
    assert np.array_equal(s1.aabb(), obs.aabb())
    assert obs.subdivide(0).counts() == obs.counts()
    assert obs.subdivide(2).n_t == 16 * obs.n_t   # BASELINE config 5, "16x-subdivided"


def test_sharding_partition():
    from robocheck_b200 import sharding as sh
    for n in (0, 1, 31, 32, 33, 1000, 100000, 1_000_000):
        for world in (1, 2, 4, 8):
            seen = 0
            for r in range(world):
                lo, hi, per = sh.shard_bounds(n, world, r)
                assert per % 32 == 0 and lo == min(n, r * per) and 0 <= hi - lo <= per
                seen += hi - lo
            assert seen == n
    flags = (np.arange(77) % 3 == 0).astype(np.uint8)
    w = sh.pack_bitmask(flags, 3)
    assert w.dtype == np.uint32 and len(w) == 3
    assert np.array_equal(sh.unpack_bitmask(w, 77), flags)
    assert (int(w[0]) >> 3) & 1 == 1 and (int(w[0]) >> 4) & 1 == 0


def test_shard_size_and_multi_argument_checks(rc):
    """rc_shard_size: shards of whole bitmask words that cover n; rc_check_multi rejects
    bad arguments with a status instead of touching a device."""
    lib = rc.lib()
    for n, g in ((0, 1), (1, 1), (100_000, 8), (10_000_000, 8), (1_000_000, 3), (33, 4), (31, 2)):
        per = lib.rc_shard_size(n, g)
        assert per % 32 == 0 and per * g >= n and (per == 0) == (n == 0)
        if n:
            assert (per - 32) * g < n + 32 * g   # no more than one word of slack per shard
    assert lib.rc_shard_size(100_000, 8) == 12512
    assert lib.rc_shard_size(5, 0) == rc.RC_ERR_INVALID and lib.rc_shard_size(-1, 2) == rc.RC_ERR_INVALID
    assert lib.rc_check_multi(None, 2, None, 10, None, None, None) == rc.RC_ERR_INVALID
    import ctypes
    handles = (ctypes.c_void_p * 2)(None, None)
    confs = np.zeros((4, 6), dtype=np.float32)
    mask = np.zeros(1, dtype=np.uint32)
    assert lib.rc_check_multi(handles, 2, confs.ctypes.data, 4, mask.ctypes.data, None, None) == rc.RC_ERR_INVALID


def test_gather_and_coplanar_argument_checks(rc):
    """rc_gather_buffer_words sizes the peer buffer (two gather areas + a signal row + a four-word
    header: call number, timeouts); rc_check_gather and rc_set_coplanar_mode reject bad arguments with
    a status, without touching a device."""
    import ctypes
    lib = rc.lib()
    assert lib.rc_gather_buffer_words(8, 3125) == 2 * 8 * 3125 + 8 + 4
    assert lib.rc_gather_buffer_words(1, 1) == 3 + 4
    for world, words in ((0, 10), (17, 10), (2, 0), (-1, 5)):
        assert lib.rc_gather_buffer_words(world, words) == rc.RC_ERR_INVALID
    peers = (ctypes.c_void_p * 2)(None, None)
    res = ctypes.c_void_p()
    assert lib.rc_check_gather(None, None, 0, None, peers, 0, 2, 10, ctypes.byref(res), None) == rc.RC_ERR_INVALID
    assert lib.rc_check_gather_begin(None, None, 0, None, peers, 0, 2, 10, ctypes.byref(res), None) == rc.RC_ERR_INVALID
    assert lib.rc_gather_end(None, peers, 0, 2, 10, None) == rc.RC_ERR_INVALID
    assert lib.rc_gather_status(None, None, 2, 10, None, None) == rc.RC_ERR_INVALID
    # audits and device known-answer tests: argument checks come before any device work
    assert lib.rc_hit_counts(None, None, 4, rc.RC_MODE_BRUTE, None, None, None) == rc.RC_ERR_INVALID
    assert lib.rc_audit(None, None, 4, None, None, None, 0) == rc.RC_ERR_INVALID
    assert lib.rc_kat_pairs(None, None, None, 1, None, None) == rc.RC_ERR_INVALID
    assert lib.rc_kat_la_solve(None, None, 1, None) == rc.RC_ERR_INVALID
    assert lib.rc_check(None, None, 4, None, None, 7, None) == rc.RC_ERR_INVALID
    assert lib.rc_check_host(None, None, 4, None, 7) == rc.RC_ERR_INVALID
    assert lib.rc_set_coplanar_mode(None, rc.RC_COPLANAR_DECIDE) == rc.RC_ERR_INVALID
    assert lib.rc_coplanar_mode(None) == rc.RC_ERR_INVALID
    assert (rc.RC_COPLANAR_SKIP, rc.RC_COPLANAR_DECIDE) == (0, 1)


def test_host_io_properties(rc, tmp_path):
    """Property checks of the loaders on generated input (hypothesis): polygons of any arity in any of
    the OBJ token forms are fanned around their first corner (Utils_rai.cu:139-320), and a .conf file
    written from what was read from a .conf file is read back bit for bit (6-digit text, :25-34)."""
    from hypothesis import given, settings, strategies as st

    coord = st.floats(min_value=-1e3, max_value=1e3, allow_nan=False, width=32)

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.tuples(coord, coord, coord), min_size=3, max_size=24),
           st.lists(st.tuples(st.integers(3, 7), st.integers(0, 2)), min_size=1, max_size=12), st.randoms())
    def obj_case(verts, faces, rnd):
        lines = ["# generated"] + [f"v {x!r} {y!r} {z!r}" for x, y, z in verts] + ["vn 0 0 1", "vt 0 0"]
        want = []
        for arity, form in faces:
            idx = [rnd.randrange(len(verts)) for _ in range(arity)]
            tok = {0: lambda i: f"{i + 1}", 1: lambda i: f"{i + 1}//1", 2: lambda i: f"{i + 1}/1/1"}[form]
            lines.append("f " + " ".join(tok(i) for i in idx))
            want += [(idx[0], idx[k], idx[k + 1]) for k in range(1, arity - 1)]
        p = tmp_path / "h.obj"
        p.write_text("\n".join(lines) + "\n")
        x, y, z, v1, v2, v3 = rc.loadOBJFile(str(p))
        assert np.array_equal(x, np.float32([v[0] for v in verts]))
        assert np.array_equal(z, np.float32([v[2] for v in verts]))
        assert list(zip(v1.tolist(), v2.tolist(), v3.tolist())) == want

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.tuples(*([st.floats(min_value=-700, max_value=700, allow_nan=False, width=32)] * 6)),
                    min_size=0, max_size=40), st.randoms())
    def conf_case(rows, rnd):
        confs = np.array(rows, np.float32).reshape(-1, 6)
        tags = np.array([rnd.randrange(2) for _ in rows], np.uint8)
        p = str(tmp_path / "h.conf")
        rc.writeConfigurationToFileTagged(confs, tags, p)
        once, t1 = rc.readConfigurationFromFileTagged(p)
        assert once.shape == confs.shape and np.array_equal(t1, tags)
        assert np.allclose(once, confs, rtol=1e-5, atol=1e-30)      # 6 significant digits
        rc.writeConfigurationToFileTagged(once, t1, p)
        twice, t2 = rc.readConfigurationFromFileTagged(p)
        assert np.array_equal(twice.view(np.uint32), once.view(np.uint32)) and np.array_equal(t2, tags)

    obj_case()
    conf_case()


ALPHA_ENV = """Boundary Box [-200:200 ; -200:200 ; -200:200]

Multibodies
2

Active
1
robot.obj -c(0 1 0 1) Volumetric Rotational
Connections
0

Passive
obstacle.obj -c(1 1 0 1) 11.0899 9.00906 3.24427  0 0 0
"""
ALPHA_QUERY = """ 0 78.240253 24.147785 -8.133371 0.286451 0.769112 0.706202
 0 -1.393717 14.630385 0.415858 0.713232 0.156663 0.826461
"""


def test_env_and_query_readers(rc, meshes, tmp_path):
    """the benchmark's .env / .query (text of models/alpha1.0/alpha.env, alpha.query):
    the environment names robot (active) and obstacle (passive) and places the obstacle at
    11.0899 9.00906 3.24427 (:13) -- a placement the reference ignores, so it is applied only on
    request (Mesh.place) and the default context stays on the untransformed mesh."""
    env_p, q_p = tmp_path / "alpha.env", tmp_path / "alpha.query"
    env_p.write_text(ALPHA_ENV)
    q_p.write_text(ALPHA_QUERY)
    env = rc.readEnvironment(str(env_p))
    assert env["box"] == ((-200.0, -200.0, -200.0), (200.0, 200.0, 200.0))
    rob_b, obs_b = env["bodies"]
    assert rob_b["file"] == "robot.obj" and rob_b["active"] and rob_b["placement"] is None
    assert rob_b["color"] == (0.0, 1.0, 0.0, 1.0)
    assert obs_b["file"] == "obstacle.obj" and not obs_b["active"]
    assert np.allclose(obs_b["placement"], (11.0899, 9.00906, 3.24427, 0, 0, 0), rtol=1e-7)
    q = rc.readQuery(str(q_p))
    assert q.shape == (2, 6)
    assert np.allclose(q[:, :3], [[78.240253, 24.147785, -8.133371], [-1.393717, 14.630385, 0.415858]], rtol=1e-7)
    turns = np.array([[0.286451, 0.769112, 0.706202], [0.713232, 0.156663, 0.826461]])  # about x, y, z
    want = (turns * 2 * np.pi + np.pi) % (2 * np.pi) - np.pi
    assert np.allclose(q[:, 5], want[:, 0], atol=1e-5)   # roll  (x)
    assert np.allclose(q[:, 3], want[:, 1], atol=1e-5)   # pitch (y)
    assert np.allclose(q[:, 4], want[:, 2], atol=1e-5)   # yaw   (z)
    assert np.all(np.abs(q[:, 3:]) <= np.pi)
    # placement: pure translation moves every vertex and the AABB by exactly the offset ...
    obs = meshes[1]
    placed = obs.place(*obs_b["placement"])
    off = np.float32(obs_b["placement"][:3])
    assert np.allclose(placed.x, obs.x + off[0], atol=1e-5) and np.allclose(placed.z, obs.z + off[2], atol=1e-5)
    assert np.array_equal(placed.v1, obs.v1) and placed.counts() == obs.counts()
    # ... a quarter turn about z maps (x, y, z) to (-y, x, z); a zero placement is the identity
    turned = obs.place(0, 0, 0, 0, 0, 90)
    assert np.allclose(turned.x, -obs.y, atol=1e-4) and np.allclose(turned.y, obs.x, atol=1e-4)
    same = obs.place()
    assert np.array_equal(same.x, obs.x) and np.array_equal(same.y, obs.y) and np.array_equal(same.z, obs.z)
    # malformed input is an error code, never a crash
    bad = tmp_path / "bad.env"
    bad.write_text("Passive\nobstacle.obj 1 2\n")
    with pytest.raises(rc.RoboCheckError) as e:
        rc.readEnvironment(str(bad))
    assert e.value.code == rc.RC_ERR_FORMAT
    with pytest.raises(rc.RoboCheckError) as e:
        rc.readEnvironment(str(tmp_path / "missing.env"))
    assert e.value.code == rc.RC_ERR_IO
    bad_q = tmp_path / "bad.query"
    bad_q.write_text("0 1 2 3\n")
    with pytest.raises(rc.RoboCheckError):
        rc.readQuery(str(bad_q))


def test_compat_header_declares_the_baselines():
    """narrow-phase.hu:37 / broad-phase.hu:17: the reference's test_baseline call site
    (test-narrow-phase.cu:628-686) must find the names it calls."""
    text = open(os.path.join(ROOT, "include", "robocheck_compat.hpp")).read()
    for name in ("narrowPhaseBaseline(", "broadPhaseBaseline(", "narrowPhase(", "broadPhaseFused_sep(", "CallMegaKernel("):
        assert name in text, name
    import subprocess
    out = subprocess.run(["nm", "-DC", os.path.join(ROOT, "robo-check_b200", "librobocheck.so")],
                         capture_output=True, text=True).stdout
    assert "narrowPhaseBaseline(" in out and "broadPhaseBaseline(" in out
