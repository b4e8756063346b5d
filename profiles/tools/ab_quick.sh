#!/bin/bash
# quick A/B of environment knobs on the bench workload: ab_quick.sh "RC_X=0" "RC_X=1" ...   (3 interleaved repetitions)
mkdir -p gpurun_out
run() { env $1 timeout 120 python bench.py --steps 40 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-28s %.1fM configs/s  %.4f ms  e2e %.1fM' % ('$1', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))" || echo "$1 FAILED"; }
for rep in 1 2 3; do for v in "$@"; do run "$v"; done; done
