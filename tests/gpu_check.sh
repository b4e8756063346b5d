#!/bin/bash
# GPU box: parity tests, then a short bench (used through gpurun while developing)
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:---no-extras} 2> gpurun_out/bench.err | tee gpurun_out/bench_last.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1fM configs/s  ms/step %.3f  e2e %.1fM  clocks %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['clocks']))
print(d['roofline']['executed'], d['roofline']['mix'])
for k in ('roofline_brute','roofline_stage','cpu_baseline'):
    if k in d: print(k, d[k])
"
tail -3 gpurun_out/bench.err
