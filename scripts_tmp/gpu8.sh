timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 900 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo bench rc=$?
tail -3 gpurun_out/bench_r02_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_b.json').read())
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['pipelined']['value'])
print('roofline', d['roofline']['frac'], d['roofline']['executed_gflop'], d['roofline']['work_avoided_x'])
print({k:(v['value'], v['ms']) for k,v in d.items() if k.startswith('scale_')})
print(d.get('parity')); print(d.get('vs_cpu_tree')); print(d.get('cpu_baseline')); print(d['cpu_baseline_tree']['value'])
PY
