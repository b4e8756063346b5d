#!/bin/bash
export RC_LIB=$PWD/ab/librobocheck_tl.so
for L in 8 64 256 1024; do
echo "== RC_SPLIT_LOAD=$L"
RC_SPLIT_LOAD=$L timeout 200 python profiles/tools/timeline_c5.py 2>&1 | grep -E "event-timed|warp leaves|entries/warp"
done
echo "== RC_SPLIT_LOAD=1024 RC_SPLIT_MAX=16"
RC_SPLIT_LOAD=1024 RC_SPLIT_MAX=16 timeout 200 python profiles/tools/timeline_c5.py 2>&1 | grep -E "event-timed|warp leaves|entries/warp"
