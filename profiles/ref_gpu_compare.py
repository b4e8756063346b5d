#!/usr/bin/env python
"""ORACLE-GPU at scale: the reference's own CUDA pipeline (broadPhaseFused_sep, compiled unchanged
for sm_100a into oracle/_ref/libref.so) against this repo's fused kernel on the same seeded sets:
flag mismatches (must be 0) and wall time of both, same B200.  The reference call includes its own
per-call set-up (OBJ parse, cudaMalloc, constant upload), exactly as its driver measures it."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robocheck_b200 as rc
from oracle import pyoracle as po
from tests import refpipe

ctx = rc.Context(device=0)
for name, confs in (("hard 100,000 (seed 508)", rc.createAlphaBotConfigurations(100_000, True, 508)),
                    ("easy 100,000 (seed 510)", rc.createAlphaBotConfigurations(100_000, False, 510)),
                    ("hard 400,000 (seed 20261018)", rc.createAlphaBotConfigurations(400_000, True, 20261018))):
    refpipe.oracle_gpu_flags(rc, po, confs[:1000])          # warm the reference (context, module load)
    t0 = time.perf_counter()
    want = refpipe.oracle_gpu_flags(rc, po, confs)
    t_ref = time.perf_counter() - t0
    ctx.check_host(confs)                                   # staging buffers sized once (kept by the context)
    t0 = time.perf_counter()
    got = ctx.check_host(confs)
    t_new = time.perf_counter() - t0
    print(json.dumps({"set": name, "configs": len(confs), "flag_mismatches": int((got != want).sum()),
                      "valid": int(want.sum()), "reference_gpu_pipeline_s": round(t_ref, 4),
                      "reference_configs_per_s": round(len(confs) / t_ref),
                      "this_repo_rc_check_host_s": round(t_new, 5),
                      "this_repo_configs_per_s": round(len(confs) / t_new),
                      "speedup_end_to_end": round(t_ref / t_new, 1)}), flush=True)
