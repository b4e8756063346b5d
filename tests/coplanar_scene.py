"""Scene for the coplanar-pair option (RC_COPLANAR_DECIDE).

A robot of a few large triangles in its local plane z = 0, an obstacle of triangles scattered over
ONE world plane, and configurations that all share one rotation and translate the robot inside that
plane (the plane is the image of z = 0 under the shared rotation, lifted off the origin so that the
plane offsets d keep a stable ratio).  Every (robot triangle, obstacle triangle) pair is then
coplanar up to float32 rounding: for a good share of them the rounding noise makes both plane
tests pass while is_coplanar_sep (narrow-phase.cu:297-338) calls the planes proportional -- the
case the reference kernel only prints an error for (:680).  Measured with the oracle (seed 6):
3,000 configurations, 1,131 such pairs, 269 configurations change their flag between the modes.

Returns (robot arrays, obstacle arrays, configurations); arrays are (x, y, z, v1, v2, v3)."""
import numpy as np


def build(po, n_rob_t=2, n_obs_t=48, n_confs=3000, seed=6, spread=80.0, size=15.0, height=25.0):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(-np.pi, np.pi, size=3).astype(np.float32)  # pitch, yaw, roll of every configuration
    rv = np.zeros((3 * n_rob_t, 3), np.float32)
    rv[:, :2] = (rng.uniform(-12, 12, size=(n_rob_t, 1, 2)) +
                 rng.uniform(-size, size, size=(n_rob_t, 3, 2))).reshape(-1, 2)
    rt = np.arange(3 * n_rob_t, dtype=np.int32).reshape(-1, 3)
    # world images of the local axes under the shared rotation, from the oracle's float32 matrix
    m = po.matrix(np.array([0, 0, 0, *ang], np.float32)).astype(np.float64).reshape(3, 4)
    u, v, w = m[:, 0], m[:, 1], m[:, 2]
    off = height * w
    ab = rng.uniform(-spread, spread, size=(n_obs_t, 1, 2)) + rng.uniform(-size, size, size=(n_obs_t, 3, 2))
    ov = (off + ab[..., :1] * u + ab[..., 1:] * v).astype(np.float32).reshape(-1, 3)
    ot = np.arange(3 * n_obs_t, dtype=np.int32).reshape(-1, 3)
    st = rng.uniform(-spread, spread, size=(n_confs, 2))
    confs = np.zeros((n_confs, 6), np.float32)
    confs[:, :3] = (off + st[:, :1] * u + st[:, 1:] * v).astype(np.float32)
    confs[:, 3:] = ang

    def cols(a):
        return tuple(a[:, i].copy() for i in range(3))

    return cols(rv) + cols(rt), cols(ov) + cols(ot), confs
