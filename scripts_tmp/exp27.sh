#!/bin/bash
for c in 4 8 12 16 4 8; do
RC_FEED_CLS=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-scale 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cls $c  e2e %.1fM %s' % (d['e2e']['value']/1e6, [round(x/1e6,1) for x in d['e2e']['runs_configs_per_s']]))"
done
