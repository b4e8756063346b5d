for g in 8 -1 8 -1; do RC_GMAIL=$g python bench.py --no-extras --no-scale 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gmail $g', 'value %.1fM %.4f ms' % (d['value']/1e6, d['ms_per_step']), 'e2e %.1fM' % (d['e2e']['value']/1e6), 'pipelined %.1fM' % (d['e2e']['pipelined']['value']/1e6))"; done
python profiles/tools/quick.py cur 2>&1 | head -4
python bench.py --no-extras --no-scale --preheat 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('preheat 0', 'value %.1fM %.4f ms' % (d['value']/1e6, d['ms_per_step']))"
