#!/bin/bash
timeout 1200 python - <<'PY'
import subprocess, sys, os, json
code = open("scripts_tmp/c5knobs.py").read().split("'''")[1]
open("/tmp/c5k.py", "w").write(code)
for env in ({"RC_GMAIL": "-1"}, {"RC_GMAIL": "-1", "RC_DONATE_MIN": "32"}, {"RC_GMAIL": "-1", "RC_DONATE_MIN": "24"}, {"RC_GMAIL": "-1", "RC_DONATE_MIN": "16"},
            {"RC_GMAIL": "-1", "RC_DONATE_MIN": "8"}, {"RC_DONATE_MIN": "16"}, {"RC_DONATE_MIN": "8"}, {"RC_DONATE_MIN": "16", "RC_MAIL_SLOTS": "8"},
            {"RC_GMAIL": "-1", "RC_DONATE_MIN": "16", "RC_DONATE_LEAF": "8"}):
    e = dict(os.environ); e.update(env); e["TAG"] = json.dumps(env)
    subprocess.run([sys.executable, "/tmp/c5k.py"], env=e, timeout=200)
PY
