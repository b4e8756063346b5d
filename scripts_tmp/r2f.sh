#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 270 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; echo rc=$?
tail -2 gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['config']['collective_check'])
print({k:(v['value'], v['ms'], v['collective_check']) for k,v in d.items() if k.startswith('scale_')})
PY
