#!/bin/bash
V=$PWD/ab/librobocheck_v5.so
export RC_LIB=$V
echo "== correctness, RC_FEED=2 (every culled check through the single-launch form)"
RC_FEED=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "culled_matches_oracle or check_host_matches or empty_and_ragged or culled_equals_brute_large or golden or million or planner_sized or pipelined or full_size_workloads or coplanar_mode or scheduling_variants or different_streams" 2>&1 | tail -5
echo "== correctness, default (host path only)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "check_host_matches or empty_and_ragged or cpp_harness or pipelined" 2>&1 | tail -3
echo "== device-resident"
RC_FEED=0 timeout 300 python profiles/tools/small.py two_launches
RC_FEED=2 timeout 300 python profiles/tools/small.py single_launch
RC_FEED=2 RC_FEED_CLS=2 timeout 300 python profiles/tools/small.py single_launch_cls2
RC_FEED=2 RC_FEED_CLS=32 timeout 300 python profiles/tools/small.py single_launch_cls32
echo "== bench, host path"
for f in 0 1; do
RC_FEED=$f timeout 300 python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('RC_FEED=$f value %.1fM  %.4f ms  e2e %.1fM %s pipelined %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, [round(x/1e6,1) for x in d['e2e']['runs_configs_per_s']], d['e2e']['pipelined']['value']/1e6))"
done
RC_FEED=1 RC_FEED_CLS=4 timeout 300 python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('RC_FEED=1 cls4 value %.1fM  %.4f ms  e2e %.1fM %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, [round(x/1e6,1) for x in d['e2e']['runs_configs_per_s']]))"
RC_FEED=1 RC_FEED_CLS=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('RC_FEED=1 cls1 value %.1fM  %.4f ms  e2e %.1fM %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, [round(x/1e6,1) for x in d['e2e']['runs_configs_per_s']]))"
