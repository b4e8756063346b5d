#!/bin/bash
# run_workloads.py under the given environment, one line per workload:  wl.sh NAME [ENV=..]...
name=$1; shift
env "$@" timeout 400 python profiles/run_workloads.py 2>gpurun_out/workloads.err | tee gpurun_out/workloads_$name.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$name  %-70s %9.4f ms %8.1f M/s' % (d['workload'][:70], d['ms'], d['configs_per_s']/1e6))"
