#!/usr/bin/env python
"""Run on a GPU box (gpurun): flags and a vertex checksum produced by the REFERENCE's own CUDA
pipeline (oracle/_ref/libref.so = the reference's kernels compiled unchanged for sm_100a) on a
seeded configuration set.  The result is committed as tests/golden/refgpu_hard.npz so that the CPU
suite can pin the oracle's GPU variant against the reference GPU path without a GPU.

    gpurun -- python tests/golden/make_refgpu_fixture.py     # writes gpurun_out/refgpu_hard.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import robocheck_b200 as rc
from oracle import pyoracle as po
from tests import refpipe

confs = np.concatenate([rc.createAlphaBotConfigurations(4000, True, 20261018),
                        rc.createAlphaBotConfigurations(1000, False, 20261019)]).astype(np.float32)
flags = refpipe.oracle_gpu_flags(rc, po, confs)
rob_m, obs_m = rc.alpha_meshes()
rob = po.OracleMesh(*rob_m.arrays())
obs_box = po.aabb(*obs_m.arrays()[:3])
ox, oy, oz, broad = po.ref_gpu_transform(rob, obs_box, confs[:64])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "refgpu_hard.npz"), confs=confs, flags=flags,
                    vx=ox, vy=oy, vz=oz, broad=broad)
print("refgpu_hard:", len(confs), "configs,", int(flags.sum()), "valid")
