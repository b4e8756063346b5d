set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
bash profiles/tools/ab_libs.sh ab/librobocheck_r01.so ab/librobocheck_r02b.so > gpurun_out/ab2.txt 2>&1
cat gpurun_out/ab2.txt
