#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_launch or check_host or pipelined or cpp_harness or ragged or different_streams" 2>&1 | tail -3
timeout 200 python profiles/tools/host_small.py final
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-scale 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1fM  %.4f ms  e2e %.1fM %s pipelined %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, [round(x/1e6,1) for x in d['e2e']['runs_configs_per_s']], d['e2e']['pipelined']['value']/1e6))"
