RC_GMAIL=-1 bash profiles/tools/ab_libs.sh ab/librobocheck_r02g.so ab/librobocheck_r02h.so ab/librobocheck_r02i.so > gpurun_out/ab12.txt 2>&1
cat gpurun_out/ab12.txt
