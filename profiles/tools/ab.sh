#!/bin/bash
# generic A/B on the bench: ab.sh "ENV1=a ENV2=b" "ENV1=c" ...   (each variant run twice, interleaved)
mkdir -p gpurun_out
run() { env $1 timeout 150 python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-28s %.1fM configs/s  %.4f ms  e2e %.1fM' % ('$1', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))" || echo "$1 FAILED"; }
for rep in 1 2; do for v in "$@"; do run "$v"; done; done
