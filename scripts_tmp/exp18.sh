RC_FEED=2 RC_LIB=$PWD/ab/librobocheck_tl.so timeout 200 python profiles/tools/timeline_feed.py 100000
