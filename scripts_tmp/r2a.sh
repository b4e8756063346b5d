#!/bin/bash
# round-2 evidence pass A: GPU suite, default bench line, launch list
mkdir -p gpurun_out
nvidia-smi -L
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err; echo bench rc=$?
tail -4 gpurun_out/bench_r02_1gpu.err
( time python bench.py --impl reference ) > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; echo ref rc=$?
bash profiles/collect.sh launches
python __graft_entry__.py smoke 2>&1 | tail -2
