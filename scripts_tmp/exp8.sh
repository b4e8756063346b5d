#!/bin/bash
echo "== correctness (mail between CTAs on for the global-obstacle kernels)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "subdivided or global_obstacle or c5_level5 or stack_overflow" 2>&1 | tail -3
echo "== timeline, one C5 shard"
for G in -1 -2; do
echo "-- RC_GMAIL=$G"
RC_GMAIL=$G RC_LIB=$PWD/ab/librobocheck_tl.so timeout 200 python profiles/tools/timeline_c5.py 2>&1 | grep -E "event-timed|warp leaves|entries/warp|not yet gone"
done
echo "== production build, C5 timings (quick.py --c5)"
for G in -1 -2 -1 -2; do
RC_GMAIL=$G timeout 300 python profiles/tools/quick.py gmail$G --c5 2>&1 | grep -E "c5|hard100k\"" | cut -c1-160
done
