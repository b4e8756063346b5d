#!/bin/bash
mkdir -p gpurun_out
timeout 60 python profiles/tools/hangcheck.py 3000 1000 | tail -1 || { echo "hangcheck failed"; exit 1; }
run() { env $1 timeout 150 python bench.py --steps 30 --warmup 5 --no-extras 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-40s %.1fM configs/s  %.4f ms  e2e %.1fM' % ('$1', d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))" || { echo "$1 FAILED"; tail -3 gpurun_out/sweep.err; }; }
{
for v in "$@"; do run "$v"; done
for v in "$@"; do run "$v"; done
} 2>&1 | tee gpurun_out/sweep.txt
