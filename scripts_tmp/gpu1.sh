set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
QUICK_ARGS=--c5 bash profiles/tools/ab_libs.sh ab/librobocheck_r01.so ab/librobocheck_r02a.so > gpurun_out/ab1.txt 2>&1
cat gpurun_out/ab1.txt
timeout 600 python bench.py > gpurun_out/bench_r02_first.json 2> gpurun_out/bench_r02_first.err; echo bench rc=$?
tail -5 gpurun_out/bench_r02_first.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_first.json').read())
print({k:d[k] for k in ('value','ms_per_step','clocks')}, d['e2e']['value'], d['e2e']['pipelined']['value'])
print(d['roofline']['frac'], d['roofline']['executed_gflop'], d['roofline']['work_avoided_x'])
print(d.get('scale_c4'), d.get('scale_c5'))
print(d.get('parity'))
PY
