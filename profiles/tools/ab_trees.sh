#!/bin/bash
# A/B of two source trees on the same box: ab_trees.sh DIR_A DIR_B   (bench workload, 3 interleaved repetitions)
run() { (cd $1 && timeout 120 python bench.py --steps 40 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-12s %.1fM configs/s  %.4f ms  e2e %.1fM' % ('$1', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))" || echo "$1 FAILED"); }
for rep in 1 2 3; do for v in "$@"; do run "$v"; done; done
