import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "robo-check_b200", "librobocheck.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "robo-check_b200"), "--no-print-directory"])
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "--no-print-directory",
                               os.path.join(ROOT, "oracle", "liboracle.so")])


_ensure_built()


@pytest.fixture(scope="session")
def rc():
    import robocheck_b200
    return robocheck_b200


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def meshes(rc):
    return rc.alpha_meshes()


@pytest.fixture(scope="session")
def omeshes(meshes, po):
    rob, obs = meshes
    return po.OracleMesh(*rob.arrays()), po.OracleMesh(*obs.arrays())


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {k: np.load(os.path.join(d, k + ".npz")) for k in
            ("easy_confs10k", "hard_confs10k", "refcpu_hard", "refgpu_hard")}


@pytest.fixture(scope="session")
def ctx(rc, meshes):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = rc.Context(meshes[0], meshes[1], device=0)
    yield c
    c.close()
