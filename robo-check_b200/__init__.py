"""robocheck_b200 -- Python host mirror of the C ABI in include/robocheck.h.

The product is `librobocheck.so` (hand-written sm_100a CUDA + C++ host code).  This module is
the thin ctypes layer the tests, bench.py and Python callers use; names follow the reference's
own entry points (broadPhaseFused_sep, CallMegaKernel, generateAABB, broadPhase, narrowPhase,
loadOBJFile, readConfigurationFromFile, createAlphaBotConfigurations ...).

There is no CPU fallback: a missing library raises ImportError here, and without a CUDA
device `Context(...)` raises `RoboCheckError(RC_ERR_NO_DEVICE)`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
# RC_LIB: another build of the same library (A/B runs of two source trees in profiles/tools)
LIB_PATH = os.environ.get("RC_LIB") or os.path.join(_HERE, "librobocheck.so")
ASSET_DIR = os.path.join(REPO_ROOT, "assets", "alpha1.0")

RC_OK = 0
RC_ERR_INVALID, RC_ERR_CUDA, RC_ERR_NOMEM, RC_ERR_IO = -1, -2, -3, -4
RC_ERR_FORMAT, RC_ERR_UNSUPPORTED, RC_ERR_NO_DEVICE = -5, -6, -7
RC_MODE_CULLED, RC_MODE_BRUTE = 0, 1
RC_COPLANAR_SKIP, RC_COPLANAR_DECIDE = 0, 1

# Configuration = 6 x f32 (x, y, z, pitch, yaw, roll), Utils_rai.h:51-58
CONFIG_DTYPE = np.dtype(np.float32)
AABB_FIELDS = ("x_min", "y_min", "z_min", "x_max", "y_max", "z_max")


class RoboCheckError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = _lib.rc_strerror(code).decode() if _lib is not None else str(code)
        super().__init__(f"robocheck error {code}: {msg}" + (f" ({detail})" if detail else ""))


RC_ENV_MAX_BODIES = 16


class rc_env_body(C.Structure):
    _fields_ = [("file", C.c_char * 256), ("active", C.c_int), ("color", C.c_float * 4),
                ("has_placement", C.c_int), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float),
                ("rx", C.c_float), ("ry", C.c_float), ("rz", C.c_float)]


class rc_env(C.Structure):
    _fields_ = [("has_box", C.c_int), ("box_lo", C.c_float * 3), ("box_hi", C.c_float * 3),
                ("n_bodies", C.c_int), ("bodies", rc_env_body * RC_ENV_MAX_BODIES)]


class rc_stats(C.Structure):
    _fields_ = [(k, C.c_ulonglong) for k in (
        "broad_cleared", "sphere_cleared", "narrow_valid", "narrow_invalid", "node_pairs",
        "leaf_pairs", "tri_pairs", "tail_pairs", "overflow", "launches", "node_rounds",
        "leaf_rounds", "tail_rounds", "shared_pieces", "watchdog", "watchdog_where", "prefilter_pairs",
        "plane1_pairs")]


# every symbol include/robocheck.h declares: (name, restype, argtypes)
_vp, _i, _l, _u, _f = C.c_void_p, C.c_int, C.c_long, C.c_uint, C.c_float
_pp = C.POINTER(C.c_void_p)
ABI = [
    ("rc_create", _i, [_pp] + [_vp] * 3 + [_i] + [_vp] * 3 + [_i] + [_vp] * 3 + [_i] + [_vp] * 3 + [_i, _i]),
    ("rc_create_from_meshes", _i, [_pp, _vp, _vp, _i]),
    ("rc_destroy", None, [_vp]),
    ("rc_strerror", C.c_char_p, [_i]),
    ("rc_last_error", C.c_char_p, [_vp]),
    ("rc_device", _i, [_vp]),
    ("rc_mesh_sizes", _i, [_vp] + [C.POINTER(_i)] * 4),
    ("rc_obstacle_aabb", _i, [_vp, _vp]),
    ("rc_check", _i, [_vp, _vp, _l, _vp, _vp, _i, _vp]),
    ("rc_check_host", _i, [_vp, _vp, _l, _vp, _i]),
    ("rc_shard_size", _i, [_l, _i]),
    ("rc_device_count", _i, [C.POINTER(_i)]),
    ("rc_check_multi", _i, [_vp, _i, _vp, _l, _vp, _vp, _vp]),
    ("rc_check_host_submit", _i, [_vp, _vp, _l, _vp, _i, C.POINTER(_i)]),
    ("rc_check_host_wait", _i, [_vp, _i]),
    ("rc_gather_buffer_words", _l, [_i, _l]),
    ("rc_check_gather", _i, [_vp, _vp, _l, _vp, _vp, _i, _i, _l, _vp, _vp]),
    ("rc_check_gather_begin", _i, [_vp, _vp, _l, _vp, _vp, _i, _i, _l, _vp, _vp]),
    ("rc_gather_end", _i, [_vp, _vp, _i, _i, _l, _vp]),
    ("rc_gather_status", _i, [_vp, _vp, _i, _l, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("rc_hit_counts", _i, [_vp, _vp, _l, _i, _vp, _vp, _vp]),
    ("rc_audit", _i, [_vp, _vp, _l, C.POINTER(_l), C.POINTER(_l), _vp, _l]),
    ("rc_kat_pairs", _i, [_vp, _vp, _vp, _i, _vp, _vp]),
    ("rc_kat_la_solve", _i, [_vp, _vp, _i, _vp]),
    ("rc_transform_aos", _i, [_vp, _vp, _l, _vp, _i, _vp, _vp]),
    ("rc_generate_aabb", _i, [_vp, _vp, _u, _u, _vp, _vp]),
    ("rc_generate_aabb_host", _i, [_vp, _vp, _u, _u, _vp]),
    ("rc_generate_aabb_soa", _i, [_vp, _vp, _vp, _vp, _u, _u, _vp, _vp]),
    ("rc_broad_phase", _i, [_vp, _l, _vp, _vp, _vp, _vp]),
    ("rc_broad_fused", _i, [_vp, _vp, _l, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("rc_narrow_phase", _i, [_vp, _vp, _vp, _vp, _l, _vp, _vp]),
    ("rc_narrow_phase_host", _i, [_i] * 5 + [_vp] * 5 + [_i]),
    ("rc_megakernel", _i, [_vp, _vp, _vp, _vp, _vp, _i, _vp]),
    ("rc_obstacle_tri_aabbs", _i, [_vp, _vp]),
    ("rc_conf_generate_device", _i, [_vp, _vp, _l] + [_f] * 6 + [C.c_uint64, _vp]),
    ("rc_conf_generate_device_mix", _i, [_vp, _vp, _l, _l, _vp, _vp, _i, _i, C.c_uint64, _vp]),
    ("rc_alpha_box", _i, [_i, _vp]),
    ("rc_compact", _i, [_vp, _vp, _vp, _l, _i, _vp, _vp, _vp]),
    ("rc_set_robot_spheres", _i, [_vp, _vp, _i]),
    ("rc_robot_spheres", _i, [_vp, _vp, C.POINTER(_i)]),
    ("rc_check_spheres", _i, [_vp, _vp, _l, _vp, _vp, _vp]),
    ("rc_set_coplanar_mode", _i, [_vp, _i]),
    ("rc_coplanar_mode", _i, [_vp]),
    ("rc_stats_enable", _i, [_vp, _i]),
    ("rc_stats_read", _i, [_vp, C.POINTER(rc_stats), _i]),
    ("rc_stats_trace", _i, [_vp, _vp]),
    ("rc_last_kernel_ms", _f, [_vp]),
    ("rc_measure_fp32_peak", _i, [_vp, C.POINTER(C.c_double)]),
    ("rc_mesh_load_obj", _i, [C.c_char_p, _pp]),
    ("rc_mesh_load_byu", _i, [C.c_char_p, _pp]),
    ("rc_mesh_from_arrays", _i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _pp]),
    ("rc_mesh_save_obj", _i, [_vp, C.c_char_p]),
    ("rc_mesh_save_byu", _i, [_vp, C.c_char_p]),
    ("rc_mesh_counts", _i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    ("rc_mesh_x", _vp, [_vp]), ("rc_mesh_y", _vp, [_vp]), ("rc_mesh_z", _vp, [_vp]),
    ("rc_mesh_v1", _vp, [_vp]), ("rc_mesh_v2", _vp, [_vp]), ("rc_mesh_v3", _vp, [_vp]),
    ("rc_mesh_subdivide", _i, [_vp, _i, _pp]),
    ("rc_mesh_aabb", _i, [_vp, _vp]),
    ("rc_mesh_free", None, [_vp]),
    ("rc_env_read", _i, [C.c_char_p, _vp]),
    ("rc_query_read", _i, [C.c_char_p, _pp, C.POINTER(_l)]),
    ("rc_mesh_place", _i, [_vp] + [_f] * 6 + [_pp]),
    ("rc_conf_read", _i, [C.c_char_p, _pp, _pp, C.POINTER(_l)]),
    ("rc_conf_write", _i, [C.c_char_p, _vp, _vp, _l]),
    ("rc_free", None, [_vp]),
    ("rc_conf_generate", _i, [_vp, _l] + [_f] * 6 + [C.c_uint32]),
    ("rc_conf_generate_alpha", _i, [_vp, _l, _i, C.c_uint32]),
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C robo-check_b200`.  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in ABI:
        if os.environ.get("RC_LIB") and not hasattr(lib, name):
            continue  # an older build under A/B comparison (profiles/tools): newer entry points are absent
        fn = getattr(lib, name)  # AttributeError here = the library does not export the ABI
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None
_lib = _load()


def lib():
    return _lib


def _check(code, ctx=None):
    if code != RC_OK:
        detail = ""
        if ctx is not None:
            detail = _lib.rc_last_error(ctx).decode(errors="replace")
        raise RoboCheckError(code, detail)


def _ptr(a):
    """numpy array -> void* (None stays NULL); ints are taken as raw device pointers."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    if hasattr(a, "data_ptr"):  # torch tensor
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _stream_ptr(stream):
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


# ------------------------------------------------------------------------------------------------
class Mesh:
    """Triangle mesh, SoA like the reference's SoA loadOBJFile overload (Utils_rai.cu:228-320)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.rc_mesh_free(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def from_obj(cls, path):
        h = C.c_void_p()
        _check(_lib.rc_mesh_load_obj(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def from_byu(cls, path):
        h = C.c_void_p()
        _check(_lib.rc_mesh_load_byu(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, path):
        return cls.from_obj(path) if str(path).lower().endswith(".obj") else cls.from_byu(path)

    @classmethod
    def from_arrays(cls, x, y, z, v1, v2, v3):
        x, y, z = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z))
        v1, v2, v3 = (np.ascontiguousarray(a, dtype=np.int32) for a in (v1, v2, v3))
        h = C.c_void_p()
        _check(_lib.rc_mesh_from_arrays(_ptr(x), _ptr(y), _ptr(z), len(x), _ptr(v1), _ptr(v2),
                                        _ptr(v3), len(v1), C.byref(h)))
        return cls(h)

    def counts(self):
        nv, nt = C.c_int(), C.c_int()
        _check(_lib.rc_mesh_counts(self._h, C.byref(nv), C.byref(nt)))
        return nv.value, nt.value

    def _arr(self, fn, n, ctype, dtype):
        p = fn(self._h)
        if not p or n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).astype(dtype, copy=True)

    @property
    def n_v(self):
        return self.counts()[0]

    @property
    def n_t(self):
        return self.counts()[1]

    @property
    def x(self):
        return self._arr(_lib.rc_mesh_x, self.n_v, C.c_float, np.float32)

    @property
    def y(self):
        return self._arr(_lib.rc_mesh_y, self.n_v, C.c_float, np.float32)

    @property
    def z(self):
        return self._arr(_lib.rc_mesh_z, self.n_v, C.c_float, np.float32)

    @property
    def v1(self):
        return self._arr(_lib.rc_mesh_v1, self.n_t, C.c_int, np.int32)

    @property
    def v2(self):
        return self._arr(_lib.rc_mesh_v2, self.n_t, C.c_int, np.int32)

    @property
    def v3(self):
        return self._arr(_lib.rc_mesh_v3, self.n_t, C.c_int, np.int32)

    def arrays(self):
        return self.x, self.y, self.z, self.v1, self.v2, self.v3

    def aabb(self):
        out = np.zeros(6, np.float32)
        _check(_lib.rc_mesh_aabb(self._h, _ptr(out)))
        return out

    def subdivide(self, levels):
        h = C.c_void_p()
        _check(_lib.rc_mesh_subdivide(self._h, int(levels), C.byref(h)))
        return Mesh(h)

    def place(self, x=0.0, y=0.0, z=0.0, rx=0.0, ry=0.0, rz=0.0):
        """The mesh moved to its place in the world (.env placement: translation, degrees about x, y, z)."""
        h = C.c_void_p()
        _check(_lib.rc_mesh_place(self._h, float(x), float(y), float(z), float(rx), float(ry), float(rz), C.byref(h)))
        return Mesh(h)

    def save_obj(self, path):
        _check(_lib.rc_mesh_save_obj(self._h, os.fsencode(path)))

    def save_byu(self, path):
        _check(_lib.rc_mesh_save_byu(self._h, os.fsencode(path)))


def loadOBJFile(path):
    """Utils_rai.cu:228 loadOBJFile(filename, x, y, z, v1, v2, v3) -> the six SoA arrays."""
    return Mesh.from_obj(path).arrays()


def alpha_meshes():
    """The alpha1.0 puzzle shipped with this repo in BYU format (the reference hard-wires
    ROB_FILE / OBS_FILE = models/alpha1.0/{robot,obstacle}.obj, Utils_rai.h:18-19)."""
    return (Mesh.from_byu(os.path.join(ASSET_DIR, "robot.g")),
            Mesh.from_byu(os.path.join(ASSET_DIR, "obstacle.g")))


# ---- .env / .query ------------------------------------------------------------------------------
def readEnvironment(path):
    """models/alpha1.0/alpha.env -> {"box": (lo, hi) or None, "bodies": [{file, active, color, placement}]}."""
    e = rc_env()
    _check(_lib.rc_env_read(os.fsencode(path), C.byref(e)))
    bodies = []
    for k in range(e.n_bodies):
        b = e.bodies[k]
        bodies.append({"file": b.file.decode(), "active": bool(b.active), "color": tuple(b.color),
                       "placement": (b.x, b.y, b.z, b.rx, b.ry, b.rz) if b.has_placement else None})
    box = (tuple(e.box_lo), tuple(e.box_hi)) if e.has_box else None
    return {"box": box, "bodies": bodies}


def readQuery(path):
    """models/alpha1.0/alpha.query -> configurations float32 [n, 6] (x y z pitch yaw roll): start, goal."""
    pc, n = C.c_void_p(), C.c_long()
    _check(_lib.rc_query_read(os.fsencode(path), C.byref(pc), C.byref(n)))
    try:
        a = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_float)), shape=(max(n.value, 1) * 6,))
        return a[: n.value * 6].reshape(n.value, 6).copy()
    finally:
        _lib.rc_free(pc)


# ---- .conf files and configuration sets ---------------------------------------------------------
def readConfigurationFromFileTagged(path):
    """-> (confs float32 [n, 6], tags uint8 [n]); Utils_rai.cu:43-72."""
    pc, pt, n = C.c_void_p(), C.c_void_p(), C.c_long()
    _check(_lib.rc_conf_read(os.fsencode(path), C.byref(pc), C.byref(pt), C.byref(n)))
    try:
        confs = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_float)), shape=(max(n.value, 1) * 6,))
        confs = confs[: n.value * 6].reshape(n.value, 6).copy()
        tags = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_ubyte)), shape=(max(n.value, 1),))
        tags = tags[: n.value].copy()
    finally:
        _lib.rc_free(pc)
        _lib.rc_free(pt)
    return confs, tags


def readConfigurationFromFile(path):
    """Utils_rai.cu:74-97 (drops the tag column)."""
    return readConfigurationFromFileTagged(path)[0]


def writeConfigurationToFileTagged(confs, valid, path):
    """Utils_rai.cu:11-41."""
    confs = np.ascontiguousarray(confs, dtype=np.float32).reshape(-1, 6)
    valid = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    _check(_lib.rc_conf_write(os.fsencode(path), _ptr(confs), _ptr(valid), len(confs)))


def generateConfs(n, x_min, x_max, y_min, y_max, z_min, z_max, seed):
    """Utils_rai.cu:322-348 with an explicit seed."""
    out = np.zeros((n, 6), np.float32)
    _check(_lib.rc_conf_generate(_ptr(out), n, x_min, x_max, y_min, y_max, z_min, z_max, seed))
    return out


def alpha_box(hard):
    """The reference's sampling box of createAlphaBotConfigurations (x_min, x_max, y_min, ...)."""
    out = np.zeros(6, np.float32)
    _check(_lib.rc_alpha_box(1 if hard else 0, _ptr(out)))
    return out


def createAlphaBotConfigurations(n, hard, seed):
    """Utils_rai.cu:111-136 with an explicit seed."""
    out = np.zeros((n, 6), np.float32)
    _check(_lib.rc_conf_generate_alpha(_ptr(out), n, 1 if hard else 0, seed))
    return out


# ---- device context -------------------------------------------------------------------------------
class Context:
    """One robot/obstacle pair resident on one GPU.  Pointers passed to the stage methods are
    device pointers: torch CUDA tensors, objects with .data_ptr(), or raw integers."""

    def __init__(self, robot=None, obstacle=None, device=0):
        if robot is None or obstacle is None:
            robot, obstacle = alpha_meshes()
        self.robot, self.obstacle = robot, obstacle
        h = C.c_void_p()
        _check(_lib.rc_create_from_meshes(C.byref(h), robot._h, obstacle._h, int(device)))
        self._h = h
        self.n_rob_v, self.n_rob_t = robot.counts()
        self.n_obs_v, self.n_obs_t = obstacle.counts()

    def close(self):
        if getattr(self, "_h", None):
            _lib.rc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def device(self):
        return _lib.rc_device(self._h)

    def obstacle_aabb(self):
        out = np.zeros(6, np.float32)
        _check(_lib.rc_obstacle_aabb(self._h, _ptr(out)), self._h)
        return out

    # fused entry point -------------------------------------------------------------------------
    def check(self, d_confs, n, d_valid=None, d_bitmask=None, mode=RC_MODE_CULLED, stream=None):
        _check(_lib.rc_check(self._h, _ptr(d_confs), int(n), _ptr(d_valid), _ptr(d_bitmask), mode,
                             _stream_ptr(stream)), self._h)

    def check_gather(self, d_confs, n, d_valid, peer_ptrs, rank, words_per_rank, stream=None):
        """Check this rank's n configurations and all-gather the bitmask through the peer-mapped buffers
        `peer_ptrs` (device addresses of every rank's buffer, see rc_check_gather in robocheck.h).
        Returns the word offset of the gathered bitmask inside the own buffer."""
        world = len(peer_ptrs)
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        res = C.c_void_p()
        _check(_lib.rc_check_gather(self._h, _ptr(d_confs), int(n), _ptr(d_valid), arr, int(rank), world,
                                    int(words_per_rank), C.byref(res), _stream_ptr(stream)), self._h)
        return (int(res.value) - int(peer_ptrs[rank])) // 4

    def check_gather_begin(self, d_confs, n, d_valid, peer_ptrs, rank, words_per_rank, stream=None):
        """rc_check_gather_begin: like check_gather without the wait for the peers (pair with gather_end)."""
        world = len(peer_ptrs)
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        res = C.c_void_p()
        _check(_lib.rc_check_gather_begin(self._h, _ptr(d_confs), int(n), _ptr(d_valid), arr, int(rank), world,
                                          int(words_per_rank), C.byref(res), _stream_ptr(stream)), self._h)
        return (int(res.value) - int(peer_ptrs[rank])) // 4

    def gather_end(self, peer_ptrs, rank, words_per_rank, stream=None):
        world = len(peer_ptrs)
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        _check(_lib.rc_gather_end(self._h, arr, int(rank), world, int(words_per_rank), _stream_ptr(stream)), self._h)

    def gather_status(self, own_ptr, world, words_per_rank):
        calls, timeouts = C.c_uint32(), C.c_uint32()
        _check(_lib.rc_gather_status(self._h, C.c_void_p(int(own_ptr)), int(world), int(words_per_rank),
                                     C.byref(calls), C.byref(timeouts)), self._h)
        return calls.value, timeouts.value

    # audits ----------------------------------------------------------------------------------------
    def hit_counts(self, d_confs, n, d_counts, mode=RC_MODE_BRUTE, d_valid=None, stream=None):
        """rc_hit_counts: hitting pairs per configuration, no early exit (d_counts: n uint32/int32)."""
        _check(_lib.rc_hit_counts(self._h, _ptr(d_confs), int(n), mode, _ptr(d_valid), _ptr(d_counts),
                                  _stream_ptr(stream)), self._h)

    def audit(self, d_confs, n, cap=64):
        """rc_audit: (flags that differ, hit-pair counts that differ, first indices) culled vs brute."""
        nf, nc = C.c_long(), C.c_long()
        idx = np.zeros(max(cap, 1), np.uint32)
        _check(_lib.rc_audit(self._h, _ptr(d_confs), int(n), C.byref(nf), C.byref(nc), _ptr(idx), int(cap)), self._h)
        return nf.value, nc.value, idx[: min(cap, max(nf.value, nc.value))]

    def kat_pairs(self, rob9, obs9):
        """rc_kat_pairs: per-pair intermediates of the device helpers -> (float32 [n, 32], int32 [n])."""
        rob9 = np.ascontiguousarray(rob9, dtype=np.float32).reshape(-1, 9)
        obs9 = np.ascontiguousarray(obs9, dtype=np.float32).reshape(-1, 9)
        n = len(rob9)
        out, verdict = np.zeros((n, 32), np.float32), np.zeros(n, np.int32)
        _check(_lib.rc_kat_pairs(self._h, _ptr(rob9), _ptr(obs9), n, _ptr(out), _ptr(verdict)), self._h)
        return out, verdict

    def kat_la_solve(self, in6):
        in6 = np.ascontiguousarray(in6, dtype=np.float32).reshape(-1, 6)
        out = np.zeros((len(in6), 2), np.float32)
        _check(_lib.rc_kat_la_solve(self._h, _ptr(in6), len(in6), _ptr(out)), self._h)
        return out

    def check_host(self, confs, mode=RC_MODE_CULLED):
        confs = np.ascontiguousarray(confs, dtype=np.float32).reshape(-1, 6)
        valid = np.zeros(len(confs), np.uint8)
        _check(_lib.rc_check_host(self._h, _ptr(confs), len(confs), _ptr(valid), mode), self._h)
        return valid

    def check_host_submit(self, confs, valid_out, mode=RC_MODE_CULLED):
        """Queue a batch (rc_check_host_submit): `confs` float32 [n, 6] and `valid_out` uint8 [n], both
        C-contiguous numpy arrays that stay untouched until check_host_wait(ticket) returns."""
        assert confs.dtype == np.float32 and confs.flags.c_contiguous and valid_out.dtype == np.uint8
        n = confs.shape[0]
        assert valid_out.flags.c_contiguous and valid_out.size >= n
        ticket = C.c_int(-1)
        _check(_lib.rc_check_host_submit(self._h, _ptr(confs), n, _ptr(valid_out), mode, C.byref(ticket)), self._h)
        if not hasattr(self, "_inflight"):
            self._inflight = {}
        self._inflight[ticket.value] = (confs, valid_out)  # keep the buffers alive
        return ticket.value

    def check_host_wait(self, ticket):
        _check(_lib.rc_check_host_wait(self._h, int(ticket)), self._h)
        return getattr(self, "_inflight", {}).pop(ticket, (None, None))[1]

    # stage API -----------------------------------------------------------------------------------
    def transform_aos(self, d_confs, n, d_out, d_base=None, n_v=0, stream=None):
        _check(_lib.rc_transform_aos(self._h, _ptr(d_confs), int(n), _ptr(d_base), int(n_v),
                                     _ptr(d_out), _stream_ptr(stream)), self._h)

    def generate_aabb(self, d_verts, n_v, n, d_out, stream=None):
        _check(_lib.rc_generate_aabb(self._h, _ptr(d_verts), int(n_v), int(n), _ptr(d_out),
                                     _stream_ptr(stream)), self._h)

    def generate_aabb_host(self, verts, n_v, n):
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        out = np.zeros((n, 6), np.float32)
        _check(_lib.rc_generate_aabb_host(self._h, _ptr(verts), int(n_v), int(n), _ptr(out)), self._h)
        return out

    def generate_aabb_soa(self, d_x, d_y, d_z, n_v, n, d_out, stream=None):
        _check(_lib.rc_generate_aabb_soa(self._h, _ptr(d_x), _ptr(d_y), _ptr(d_z), int(n_v), int(n),
                                         _ptr(d_out), _stream_ptr(stream)), self._h)

    def broad_phase(self, n, d_robots, d_valid, d_obstacle=None, stream=None):
        _check(_lib.rc_broad_phase(self._h, int(n), _ptr(d_robots), _ptr(d_obstacle), _ptr(d_valid),
                                   _stream_ptr(stream)), self._h)

    def broad_fused(self, d_confs, n, d_x=None, d_y=None, d_z=None, d_aabb=None, d_valid=None,
                    stream=None):
        _check(_lib.rc_broad_fused(self._h, _ptr(d_confs), int(n), _ptr(d_x), _ptr(d_y), _ptr(d_z),
                                   _ptr(d_aabb), _ptr(d_valid), _stream_ptr(stream)), self._h)

    def narrow_phase(self, d_x, d_y, d_z, n, d_valid, stream=None):
        _check(_lib.rc_narrow_phase(self._h, _ptr(d_x), _ptr(d_y), _ptr(d_z), int(n), _ptr(d_valid),
                                    _stream_ptr(stream)), self._h)

    def megakernel(self, d_configs, d_valid_confs, n, d_obs_aabb=None, d_obs_tri_aabbs=None, stream=None):
        """test/MegaKernel.hu:6 argument list (obstacle AABB arguments are accepted and ignored)."""
        _check(_lib.rc_megakernel(self._h, _ptr(d_configs), _ptr(d_obs_aabb), _ptr(d_obs_tri_aabbs),
                                  _ptr(d_valid_confs), int(n), _stream_ptr(stream)), self._h)

    def obstacle_tri_aabbs(self):
        out = np.zeros((self.n_obs_t, 6), np.float32)
        _check(_lib.rc_obstacle_tri_aabbs(self._h, _ptr(out)), self._h)
        return out

    def generate_device(self, d_out, n, box, seed, stream=None):
        _check(_lib.rc_conf_generate_device(self._h, _ptr(d_out), int(n), *[float(v) for v in box],
                                            int(seed), _stream_ptr(stream)), self._h)

    def generate_device_mix(self, d_out, first, n, box_a, box_b=None, block=1024, period=0, seed=0, stream=None):
        """rc_conf_generate_device_mix: configurations first .. first + n - 1 of one global set."""
        a = np.ascontiguousarray(box_a, dtype=np.float32)
        b = np.ascontiguousarray(box_b if box_b is not None else box_a, dtype=np.float32)
        _check(_lib.rc_conf_generate_device_mix(self._h, _ptr(d_out), int(first), int(n), _ptr(a), _ptr(b),
                                                int(block), int(period), int(seed), _stream_ptr(stream)), self._h)

    def compact(self, n, d_indices, d_count, d_valid=None, d_bitmask=None, want=1, stream=None):
        _check(_lib.rc_compact(self._h, _ptr(d_valid), _ptr(d_bitmask), int(n), int(want),
                               _ptr(d_indices), _ptr(d_count), _stream_ptr(stream)), self._h)

    # sphere-approximated robot ---------------------------------------------------------------------
    def set_robot_spheres(self, xyzr=None):
        if xyzr is None:
            _check(_lib.rc_set_robot_spheres(self._h, None, 0), self._h)
        else:
            xyzr = np.ascontiguousarray(xyzr, dtype=np.float32).reshape(-1, 4)
            _check(_lib.rc_set_robot_spheres(self._h, _ptr(xyzr), len(xyzr)), self._h)

    def robot_spheres(self):
        n = C.c_int()
        _check(_lib.rc_robot_spheres(self._h, None, C.byref(n)), self._h)
        out = np.zeros((n.value, 4), np.float32)
        if n.value:
            _check(_lib.rc_robot_spheres(self._h, _ptr(out), C.byref(n)), self._h)
        return out

    def check_spheres(self, d_confs, n, d_valid=None, d_bitmask=None, stream=None):
        _check(_lib.rc_check_spheres(self._h, _ptr(d_confs), int(n), _ptr(d_valid), _ptr(d_bitmask),
                                     _stream_ptr(stream)), self._h)

    # coplanar pairs ------------------------------------------------------------------
    def set_coplanar_mode(self, mode):
        """RC_COPLANAR_SKIP (reference behaviour, narrow-phase.cu:647-652,680) or RC_COPLANAR_DECIDE."""
        _check(_lib.rc_set_coplanar_mode(self._h, int(mode)), self._h)

    def coplanar_mode(self):
        return int(_lib.rc_coplanar_mode(self._h))

    # measurement -----------------------------------------------------------------------------------
    def stats_enable(self, on=True):
        _check(_lib.rc_stats_enable(self._h, 1 if on else 0), self._h)

    def stats_trace(self, d_trace):
        """d_trace: zero-filled device tensor of 4*n uint32/int32 words, or None to switch it off."""
        _check(_lib.rc_stats_trace(self._h, _ptr(d_trace) if d_trace is not None else None), self._h)

    def stats(self, reset=False):
        s = rc_stats()
        _check(_lib.rc_stats_read(self._h, C.byref(s), 1 if reset else 0), self._h)
        return {k: getattr(s, k) for k, _ in rc_stats._fields_}

    def last_kernel_ms(self):
        return float(_lib.rc_last_kernel_ms(self._h))

    def fp32_peak_tflops(self):
        v = C.c_double()
        _check(_lib.rc_measure_fp32_peak(self._h, C.byref(v)), self._h)
        return v.value


def check_multi(contexts, confs, want_flags=False, d_bitmask_all=None):
    """rc_check_multi: one Context per device, host configurations in, the full validity bitmask
    (and optionally the flag bytes) out; d_bitmask_all = one device tensor of (n+31)//32 int32 words
    per context to leave the full bitmask on every device as well."""
    confs = np.ascontiguousarray(confs, dtype=np.float32).reshape(-1, 6)
    n = len(confs)
    g = len(contexts)
    handles = (C.c_void_p * g)(*[c._h for c in contexts])
    mask = np.zeros((n + 31) // 32, dtype=np.uint32)
    flags = np.zeros(n, dtype=np.uint8) if want_flags else None
    dall = None
    if d_bitmask_all is not None:
        dall = (C.c_void_p * g)(*[_ptr(t) for t in d_bitmask_all])
    _check(_lib.rc_check_multi(handles, g, confs.ctypes.data, n, mask.ctypes.data,
                               flags.ctypes.data if want_flags else None, dall))
    return (mask, flags) if want_flags else mask


def narrowPhase(rob_trs, rob_pts, obs_trs, obs_pts, valid, device=0):
    """narrow-phase.hu:22 narrowPhase(): host AoS arrays, `valid` in/out ([n] uint8)."""
    rob_trs = np.ascontiguousarray(rob_trs, dtype=np.int32).reshape(-1, 3)
    obs_trs = np.ascontiguousarray(obs_trs, dtype=np.int32).reshape(-1, 3)
    rob_pts = np.ascontiguousarray(rob_pts, dtype=np.float32)
    obs_pts = np.ascontiguousarray(obs_pts, dtype=np.float32).reshape(-1, 3)
    n = len(valid)
    n_rob_v = rob_pts.size // 3 // max(n, 1)
    _check(_lib.rc_narrow_phase_host(n, len(rob_trs), n_rob_v, len(obs_trs), len(obs_pts),
                                     _ptr(rob_trs), _ptr(rob_pts), _ptr(obs_trs), _ptr(obs_pts),
                                     _ptr(valid), device))
    return valid


_default_ctx = {}


def _default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device=device)
    return _default_ctx[device]


def broadPhaseFused_sep(confs, device=0):
    """broad-phase-fused.hu:11 broadPhaseFused_sep(confs, valid_conf) on the alpha1.0 meshes."""
    return _default_context(device).check_host(confs).astype(bool)


def CallMegaKernel(confs, device=0):
    """test/MegaKernel.hu:9 CallMegaKernel(configs, valid_confs)."""
    return _default_context(device).check_host(confs).astype(bool)
