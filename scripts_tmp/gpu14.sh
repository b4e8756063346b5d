nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x -k "two_gpu or one_context_per_gpu or gather or known_answer" 2>&1 | tail -6
cat gpurun_out/fused_gather_status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; echo rc=$?
tail -3 gpurun_out/bench_r02_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_2gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['api'][:40], d['config']['collective'][:60], d['config']['collective_check'])
print({k:(v['value'], v['ms'], v['collective_check']) for k,v in d.items() if k.startswith('scale_')})
PY
