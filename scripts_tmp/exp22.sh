RC_LIB=$PWD/ab/librobocheck_tl.so timeout 200 python profiles/tools/timeline.py 2000 10000 30000 2>&1 | grep -v "last warps"
