run() { python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1  %.1fM configs/s  %.4f ms  e2e %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"; }
RC_SEED_DEPTH=0 run seed0
RC_SEED_DEPTH=1 run seed1
RC_SEED_DEPTH=2 run seed2
RC_SEED_DEPTH=3 run seed3
