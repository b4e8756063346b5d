for rep in 1 2 3; do for g in 8 -1; do RC_GMAIL=$g python profiles/tools/quick.py gmail$g 2>&1 | grep -E "hard100k|hard1M|hard10k|hard2k|easy1M" ; done; done
for rep in 1 2 3 4; do for g in 8 -1; do RC_GMAIL=$g python bench.py --no-extras --no-scale 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gmail $g', 'value %.1fM %.4f ms' % (d['value']/1e6, d['ms_per_step']))"; done; done
