#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "subdivided or global_obstacle or c5_level5 or stack_overflow" 2>&1 | tail -2
timeout 600 python - <<'PY'
import subprocess, sys, os, json
code = open("scripts_tmp/c5knobs.py").read().split("'''")[1]
open("/tmp/c5k.py", "w").write(code)
for env in ({"RC_GMAIL": "-1"}, {}, {"RC_GMAIL": "-1"}, {}):
    e = dict(os.environ); e.update(env); e["TAG"] = json.dumps(env)
    subprocess.run([sys.executable, "/tmp/c5k.py"], env=e, timeout=200)
PY
RC_LIB=$PWD/ab/librobocheck_tl.so timeout 200 python profiles/tools/timeline_c5.py 2>&1 | grep -E "event-timed|warp leaves|entries/warp|not yet gone" | head -4
