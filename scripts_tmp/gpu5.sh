timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
QUICK_ARGS=--c5 bash profiles/tools/ab_libs.sh ab/librobocheck_r02a.so ab/librobocheck_r02f.so > gpurun_out/ab5.txt 2>&1
cat gpurun_out/ab5.txt
