"""ctypes front-end of the CPU oracle (oracle/oracle.c) and of the compiled reference
(oracle/_ref/libref.so).  TEST INFRASTRUCTURE ONLY: importable from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref.so")
VARIANT_GPU, VARIANT_HOST = 0, 1
FLAG_COPLANAR = 0x10  # or-ed into a variant: decide coplanar pairs in the plane (RC_COPLANAR_DECIDE)


class rco_mesh(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p), ("v1", C.c_void_p),
                ("v2", C.c_void_p), ("v3", C.c_void_p), ("n_v", C.c_int), ("n_t", C.c_int)]


class rco_stats(C.Structure):
    _fields_ = [(k, C.c_long) for k in ("broad_cleared", "narrow_valid", "narrow_invalid",
                                        "pairs_stage1", "pairs_full", "hit_pairs", "coplanar")]


def build():
    import subprocess
    subprocess.check_call(["make", "-C", HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _lib = C.CDLL(ORACLE_SO)
        _lib.rco_project.restype = C.c_float
        _lib.rco_param.restype = C.c_float
        _lib.rco_point_tri_dist2.restype = C.c_float
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleMesh:
    def __init__(self, x, y, z, v1, v2, v3):
        self.x, self.y, self.z = _f(x), _f(y), _f(z)
        self.v1, self.v2, self.v3 = (np.ascontiguousarray(a, dtype=np.int32) for a in (v1, v2, v3))
        self.c = rco_mesh(_p(self.x), _p(self.y), _p(self.z), _p(self.v1), _p(self.v2), _p(self.v3),
                          len(self.x), len(self.v1))


def check(rob, obs, confs, variant=VARIANT_GPU, nthreads=None, early_exit=True, want_stats=False):
    """Flags (uint8, 1 = valid) for every configuration, brute force over all triangle pairs."""
    confs = _f(confs).reshape(-1, 6)
    valid = np.zeros(len(confs), np.uint8)
    st = rco_stats()
    nthreads = nthreads or os.cpu_count() or 1
    rc = lib().rco_check(C.byref(rob.c), C.byref(obs.c), _p(confs), C.c_long(len(confs)), variant,
                         nthreads, 1 if early_exit else 0, _p(valid), C.byref(st))
    assert rc == 0
    if want_stats:
        return valid, {k: getattr(st, k) for k, _ in rco_stats._fields_}
    return valid


def check_counts(rob, obs, confs, variant=VARIANT_GPU, nthreads=None):
    """(flags, hitting pairs per configuration): the exhaustive sweep without its early exit."""
    confs = _f(confs).reshape(-1, 6)
    valid = np.zeros(len(confs), np.uint8)
    counts = np.zeros(len(confs), np.uint32)
    rc = lib().rco_check_counts(C.byref(rob.c), C.byref(obs.c), _p(confs), C.c_long(len(confs)), variant,
                                nthreads or os.cpu_count() or 1, _p(valid), _p(counts))
    assert rc == 0
    return valid, counts


def pair_intermediates(r9, o9):
    """Every intermediate of the GPU variant for one pair, laid out like rc_kat_pairs' 32 floats,
    and the verdict word (plane test O | plane test R << 1 | tail << 2; tail 0 miss, 1 hit, 2 coplanar)."""
    L = lib()
    F3, F9 = C.c_float * 3, C.c_float * 9
    r, o = F9(*_f(r9)), F9(*_f(o9))
    out = np.zeros(32, np.float32)
    Nr, No, dO, dR, D, O = F3(), F3(), F3(), F3(), F3(), F3()
    dr, do_ = C.c_float(), C.c_float()
    L.rco_plane(r, 1, Nr, C.byref(dr))
    L.rco_plane(o, 1, No, C.byref(do_))
    L.rco_signed_dists(Nr, dr, o, 1, dO)
    L.rco_signed_dists(No, do_, r, 1, dR)
    out[0:3], out[3], out[4:7] = list(Nr), dr.value, list(dO)
    out[7:10], out[10], out[11:14] = list(No), do_.value, list(dR)
    sO, sR = int(not L.rco_no_overlap(dO, 1)), int(not L.rco_no_overlap(dR, 1))
    tail = 0
    if True:  # the tail is evaluated unconditionally (rc_kat_pairs does the same); it matters when sO and sR
        if L.rco_is_coplanar(Nr, dr, No, do_):
            tail = 2
        else:
            L.rco_intersect_line(Nr, dr, No, do_, 1, D, O)
            ri, oi = (C.c_int * 3)(), (C.c_int * 3)()
            L.rco_canonicalize(dR, ri)
            L.rco_canonicalize(dO, oi)
            rv = np.asarray(_f(r9)).reshape(3, 3)
            ov = np.asarray(_f(o9)).reshape(3, 3)

            def param(v, idx, d, a, b):
                return L.rco_param(F3(*v[idx[a]]), F3(*v[idx[b]]), C.c_float(d[idx[a]]), C.c_float(d[idx[b]]), D, O, 1)

            t = [param(rv, ri, dR, 0, 1), param(rv, ri, dR, 1, 2), param(ov, oi, dO, 0, 1), param(ov, oi, dO, 1, 2)]
            out[14:17], out[17:20], out[20:24] = list(D), list(O), t
            tail = 1
            with np.errstate(invalid="ignore"):
                if min(t[0], t[1]) >= max(t[2], t[3]) or min(t[2], t[3]) >= max(t[0], t[1]):
                    tail = 0
    return out, sO | (sR << 1) | (tail << 2)


def la_solve_gpu(A1, A2, A3, A4, b1, b2):
    x1, x2 = C.c_float(), C.c_float()
    lib().rco_la_solve_gpu(*(C.c_float(v) for v in (A1, A2, A3, A4, b1, b2)), C.byref(x1), C.byref(x2))
    return x1.value, x2.value


def check_bvh(rob, obs, confs, variant=VARIANT_GPU, nthreads=None, want_stats=False):
    """Same flags as check(), pairs enumerated by a dual-tree descent (oracle_bvh.c): fast enough for
    the full BASELINE sizes."""
    confs = _f(confs).reshape(-1, 6)
    valid = np.zeros(len(confs), np.uint8)
    st = rco_stats()
    nthreads = nthreads or os.cpu_count() or 1
    rc = lib().rco_check_bvh(C.byref(rob.c), C.byref(obs.c), _p(confs), C.c_long(len(confs)), variant,
                             nthreads, _p(valid), C.byref(st))
    assert rc == 0
    if want_stats:
        return valid, {k: getattr(st, k) for k, _ in rco_stats._fields_}
    return valid


def transform(confs, x, y, z):
    confs = _f(confs).reshape(-1, 6)
    x, y, z = _f(x), _f(y), _f(z)
    n, nv = len(confs), len(x)
    ox, oy, oz = (np.zeros((n, nv), np.float32) for _ in range(3))
    lib().rco_transform(_p(confs), n, _p(x), _p(y), _p(z), nv, _p(ox), _p(oy), _p(oz))
    return ox, oy, oz


def aabb(x, y, z):
    x, y, z = _f(x), _f(y), _f(z)
    out = np.zeros(6, np.float32)
    lib().rco_aabb(_p(x), _p(y), _p(z), len(x), _p(out))
    return out


def broad_valid(rob_box, obs_box):
    return int(lib().rco_broad_valid(_p(_f(rob_box)), _p(_f(obs_box))))


def matrix(conf):
    m = np.zeros(12, np.float32)
    lib().rco_matrix(_p(_f(conf)), _p(m))
    return m


def sincos(x):
    s, c = C.c_float(), C.c_float()
    lib().rco_sincosf(C.c_float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def narrow(rob, obs, px, py, pz, valid, variant=VARIANT_GPU):
    px, py, pz = _f(px), _f(py), _f(pz)
    valid = np.ascontiguousarray(valid, dtype=np.uint8).copy()
    lib().rco_narrow(C.byref(rob.c), C.byref(obs.c), _p(px), _p(py), _p(pz), C.c_long(len(valid)),
                     variant, _p(valid))
    return valid


def pair_hits(r9, o9, variant=VARIANT_GPU):
    cop = C.c_int()
    h = lib().rco_pair_hits(_p(_f(r9)), _p(_f(o9)), variant, C.byref(cop))
    return int(h), int(cop.value)


def coplanar_overlap(r9, o9, normal3):
    return int(lib().rco_coplanar_overlap(_p(_f(r9)), _p(_f(o9)), _p(_f(normal3))))


def check_spheres(spheres, obs, confs, nthreads=None):
    """Sphere-approximated robot (this repo's own definition, see oracle.c): uint8 flags."""
    spheres = _f(spheres).reshape(-1, 4)
    confs = _f(confs).reshape(-1, 6)
    valid = np.zeros(len(confs), np.uint8)
    rc = lib().rco_check_spheres(_p(spheres), len(spheres), C.byref(obs.c), _p(confs),
                                 C.c_long(len(confs)), nthreads or os.cpu_count() or 1, _p(valid))
    assert rc == 0
    return valid


def point_tri_dist2(p, tri9):
    return float(lib().rco_point_tri_dist2(_p(_f(p)), _p(_f(tri9))))


# ---- the compiled reference ------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
    return _ref


def ref_cpu_narrow(rob, obs, px, py, pz, nthreads=None):
    """narrowPhaseBaseline (reference narrow-phase.cu:342-424) on transformed SoA vertices
    [n][n_v]; converts to the AoS layout the reference expects.  Threads split the configs."""
    import threading
    n, nv = px.shape
    pts = np.ascontiguousarray(np.stack([px, py, pz], axis=-1), dtype=np.float32)  # [n][nv][3]
    rtri = np.ascontiguousarray(np.stack([rob.v1, rob.v2, rob.v3], axis=-1), dtype=np.int32)
    otri = np.ascontiguousarray(np.stack([obs.v1, obs.v2, obs.v3], axis=-1), dtype=np.int32)
    opts = np.ascontiguousarray(np.stack([obs.x, obs.y, obs.z], axis=-1), dtype=np.float32)
    flags = np.zeros(n, np.uint8)
    nthreads = max(1, min(nthreads or os.cpu_count() or 1, n))
    fn = ref().ref_cpu_narrow

    def work(lo, hi):
        if hi > lo:
            fn(hi - lo, len(rtri), nv, len(otri), len(opts), _p(rtri),
               C.c_void_p(pts.ctypes.data + lo * nv * 12), _p(otri), _p(opts),
               C.c_void_p(flags.ctypes.data + lo))

    ths = [threading.Thread(target=work, args=(n * t // nthreads, n * (t + 1) // nthreads))
           for t in range(nthreads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    return flags


def ref_gpu_check(workdir, confs):
    confs = _f(confs).reshape(-1, 6)
    flags = np.zeros(len(confs), np.uint8)
    rc = ref().ref_gpu_check(os.fsencode(workdir), _p(confs), C.c_long(len(confs)), _p(flags))
    if rc != 0:
        raise RuntimeError(f"ref_gpu_check failed: {rc}")
    return flags


def ref_gpu_transform(rob, obs_box, confs):
    confs = _f(confs).reshape(-1, 6)
    n, nv = len(confs), len(rob.x)
    ox, oy, oz = (np.zeros((n, nv), np.float32) for _ in range(3))
    flags = np.zeros(n, np.uint8)
    rc = ref().ref_gpu_transform(_p(rob.x), _p(rob.y), _p(rob.z), nv, _p(_f(obs_box)), _p(confs), n,
                                 _p(ox), _p(oy), _p(oz), _p(flags))
    if rc != 0:
        raise RuntimeError(f"ref_gpu_transform failed: {rc}")
    return ox, oy, oz, flags
