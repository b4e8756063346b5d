#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "subdivided or global_obstacle or c5_level5 or stack_overflow" 2>&1 | tail -2
timeout 300 python - <<'PY'
import subprocess, sys, os, json
code = open("scripts_tmp/c5knobs.py").read().split("'''")[1]
open("/tmp/c5k.py", "w").write(code)
for tag in ("final", "final"):
    e = dict(os.environ); e["TAG"] = tag
    subprocess.run([sys.executable, "/tmp/c5k.py"], env=e, timeout=200)
PY
