python profiles/tools/count_diff.py 2>&1 | tail -12
bash profiles/tools/ab_libs.sh ab/librobocheck_r02a.so ab/librobocheck_r02c.so ab/librobocheck_r02e.so > gpurun_out/ab4.txt 2>&1
cat gpurun_out/ab4.txt
for ch in 1 2 4; do RC_HOST_CHUNKS=$ch python bench.py --no-extras --no-scale --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunks $ch', 'value %.1fM' % (d['value']/1e6), 'e2e %.1fM' % (d['e2e']['value']/1e6), 'pipelined %.1fM' % (d['e2e']['pipelined']['value']/1e6), d['clocks'])"; done
