export RC_LIB=$PWD/ab/librobocheck_tl.so
RC_FEED=2 timeout 100 python profiles/tools/timeline.py 100000 2>&1 | grep -v "last warps\|list dry\|still drawing"
export RC_LIB=$PWD/ab/librobocheck_v2.so
for f in 0 2; do
RC_FEED=$f timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:check_culled -c 2 python scripts_tmp/dbg_feed.py 100000 hard 2>&1 | grep -E "check_culled|duration|inst_executed|issue_active|warps_active" | tail -6
done
