run() { python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1  %.1fM configs/s  %.4f ms  e2e %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"; }
run default
RC_TAIL_THRESH=24 run tail24
RC_TAIL_THRESH=16 run tail16
RC_TAIL_THRESH=8 run tail8
RC_TAIL_THRESH=4 run tail4
