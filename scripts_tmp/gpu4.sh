python profiles/tools/count_diff.py 2>&1 | tail -12
bash profiles/tools/ab_libs.sh ab/librobocheck_r02a.so ab/librobocheck_r02c.so ab/librobocheck_r02d.so > gpurun_out/ab4.txt 2>&1
cat gpurun_out/ab4.txt
