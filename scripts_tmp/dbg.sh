export RC_FEED=2 CUDA_LAUNCH_BLOCKING=1
export RC_LIB=$PWD/ab/librobocheck_v2.so
timeout 60 python scripts_tmp/dbg_feed.py 64 easy far 2>&1 | tail -1
timeout 60 python scripts_tmp/dbg_feed.py 64 hard 2>&1 | tail -1
timeout 60 python scripts_tmp/dbg_feed.py 5000 hard 2>&1 | tail -1
timeout 60 python scripts_tmp/dbg_feed.py 100000 hard 2>&1 | tail -1
