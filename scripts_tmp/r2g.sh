#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err; echo bench rc=$?
timeout 600 python profiles/run_workloads.py > gpurun_out/workloads_r02.jsonl 2> gpurun_out/workloads.err; echo workloads rc=$?
python __graft_entry__.py smoke 2>&1 | tail -1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_1gpu.json').read().strip().splitlines()[-1])
print('value %.1fM  ms %.4f  e2e %.1fM %s  pipelined %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, [round(x/1e6) for x in d['e2e']['runs_configs_per_s']], d['e2e']['pipelined']['value']/1e6))
print(d['parity']['flag_mismatches'], d['parity']['brute_flag_mismatches'], d['parity']['reference_gpu_flag_mismatches'], d['parity']['audit'])
print({k:(v['value'],v['ms']) for k,v in d.items() if k.startswith('scale_')})
for l in open('gpurun_out/workloads_r02.jsonl'):
    w=json.loads(l); print(w['workload'][:58], w['ms'], w['configs_per_s'])
PY
