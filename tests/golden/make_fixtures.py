#!/usr/bin/env python
"""Regenerates the committed fixtures from the read-only reference tree.  Run in the build
container only (needs /root/reference and oracle/_ref/libref.so):

    python tests/golden/make_fixtures.py

  assets/alpha1.0/{robot,obstacle}.g   the alpha1.0 puzzle converted OBJ -> BYU (.g), keeping the
                                       decimal text of every coordinate (so (float)atof() of the
                                       two files is identical) -- models/alpha1.0/*.obj
  tests/golden/{easy,hard}_confs10k.npz  the reference's tagged configuration sets (column 7 =
                                       FCL verdict), parsed by the product's own .conf reader
  tests/golden/refcpu_hard.npz         flags of the reference's narrowPhaseBaseline (compiled
                                       from /root/reference, run here) on the first configs of the
                                       hard set + a seeded synthetic set
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def obj_to_byu(src, dst):
    verts, faces = [], []
    for line in open(src):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            verts.append(t[1:4])
        elif t[0] == "f":
            idx = [int(tok.split("/")[0]) for tok in t[1:]]
            for k in range(len(idx) - 2):
                faces.append((idx[0], idx[k + 1], idx[k + 2]))
    with open(dst, "w") as f:
        f.write(f"1 {len(verts)} {len(faces)} {3 * len(faces)}\n1 {len(faces)}\n")
        for v in verts:
            f.write(" ".join(v) + "\n")
        for a, b, c in faces:
            f.write(f"{a} {b} -{c}\n")
    return len(verts), len(faces)


def main():
    os.makedirs(os.path.join(ROOT, "assets", "alpha1.0"), exist_ok=True)
    for name in ("robot", "obstacle"):
        nv, nt = obj_to_byu(f"{REF}/models/alpha1.0/{name}.obj", f"{ROOT}/assets/alpha1.0/{name}.g")
        print(name, nv, "vertices", nt, "triangles")

    import robocheck_b200 as rc
    from oracle import pyoracle as po

    # the BYU conversion must load to exactly the floats the reference's OBJ loader produces
    for name in ("robot", "obstacle"):
        a = rc.Mesh.from_obj(f"{REF}/models/alpha1.0/{name}.obj").arrays()
        b = rc.Mesh.from_byu(f"{ROOT}/assets/alpha1.0/{name}.g").arrays()
        assert all(np.array_equal(p, q) for p, q in zip(a, b)), name

    for kind in ("easy", "hard"):
        confs, tags = rc.readConfigurationFromFileTagged(f"{REF}/{kind}_confs10,000.conf")
        assert confs.shape == (10000, 6)
        np.savez_compressed(f"{GOLD}/{kind}_confs10k.npz", confs=confs, tags=tags)
        print(kind, "valid", int(tags.sum()), "invalid", int((1 - tags).sum()))

    # reference CPU brute force (the reference's own code, compiled unchanged, run here)
    rob_m, obs_m = rc.alpha_meshes()
    rob, obs = po.OracleMesh(*rob_m.arrays()), po.OracleMesh(*obs_m.arrays())
    hard = np.load(f"{GOLD}/hard_confs10k.npz")["confs"][:1200]
    synth = rc.createAlphaBotConfigurations(400, True, 20261018)
    confs = np.concatenate([hard, synth]).astype(np.float32)
    ox, oy, oz = po.transform(confs, rob.x, rob.y, rob.z)
    flags = po.ref_cpu_narrow(rob, obs, ox, oy, oz)
    np.savez_compressed(f"{GOLD}/refcpu_hard.npz", confs=confs, flags=flags)
    print("refcpu_hard:", len(confs), "configs,", int(flags.sum()), "valid")


if __name__ == "__main__":
    main()
