#!/bin/bash
bash profiles/collect.sh launches >/dev/null 2>&1
for n in 2000 30000 100000; do
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__block_size --clock-control none -k regex:check_culled -c 1 --launch-skip 2 python scripts_tmp/dbg_feed.py $n hard 2>&1 | grep -E "check_culled|duration|inst_executed|issue_active|warps_active|block_size" | tail -7
done
