run() { python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1  %.1fM configs/s  %.4f ms  e2e %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"; }
run default
RC_CHUNK=1 run chunk1
RC_CHUNK=2 run chunk2
RC_CHUNK=4 run chunk4
RC_CHUNK=8 run chunk8
RC_TAIL_THRESH=16 run tail16
RC_TAIL_THRESH=8 run tail8
RC_LEAF_THRESH=4 run leaf4
RC_TAIL_THRESH=16 RC_LEAF_THRESH=4 run tail16leaf4
