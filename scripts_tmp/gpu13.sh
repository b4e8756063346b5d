timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
RC_GMAIL=-1 bash profiles/tools/ab_libs.sh ab/librobocheck_r02g.so ab/librobocheck_r02j.so > gpurun_out/ab13.txt 2>&1
grep -E "hard|easy1M" gpurun_out/ab13.txt
