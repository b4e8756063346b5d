#!/bin/bash
# liveness + parity first, then knob sweep on the bench (hand-over slots / threshold, cost classes on/off)
mkdir -p gpurun_out
timeout 60 python profiles/tools/hangcheck.py 3000 1000 | tail -1 || { echo "hangcheck failed"; exit 1; }
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { env $1 timeout 150 python bench.py --steps 30 --warmup 5 --no-extras 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-40s %.1fM configs/s  %.4f ms  e2e %.1fM' % ('$1', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))" || { echo "$1 FAILED"; tail -3 gpurun_out/sweep.err; }; }
{
run "RC_BOUND=0"
run "RC_BOUND=1"
run "RC_BOUND=0"
run "RC_BOUND=1"
run "RC_BOUND=1 RC_SHARE=0"
for s in 8 16; do for d in 40 64 96; do run "RC_MAIL_SLOTS=$s RC_DONATE_MIN=$d"; done; done
run "RC_MAIL_SLOTS=4 RC_DONATE_MIN=40"
run "RC_PDL=1"
} 2>&1 | tee gpurun_out/sweep.txt
