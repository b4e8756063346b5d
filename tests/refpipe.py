"""Test-side helper around the compiled reference (oracle/_ref/libref.so): prepares the working
directory broadPhaseFused_sep() expects (./models/alpha1.0/{robot,obstacle}.obj, Utils_rai.h:18-19)
from this repo's own assets, and exposes ORACLE-GPU = the reference's CUDA pipeline on this GPU."""
import os
import tempfile

_workdir = None


def workdir(rc):
    global _workdir
    if _workdir is None:
        d = tempfile.mkdtemp(prefix="refpipe_")
        os.makedirs(os.path.join(d, "models", "alpha1.0"))
        rob, obs = rc.alpha_meshes()
        rob.save_obj(os.path.join(d, "models", "alpha1.0", "robot.obj"))
        obs.save_obj(os.path.join(d, "models", "alpha1.0", "obstacle.obj"))
        _workdir = d
    return _workdir


def oracle_gpu_flags(rc, po, confs):
    return po.ref_gpu_check(workdir(rc), confs)
