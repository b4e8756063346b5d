#!/bin/bash
# liveness of the hand-over first (tight timeout; a hung kernel wedges the box, so stop there),
# then the A/B of RC_SHARE on the bench, the per-configuration trace, and the GPU test suite
mkdir -p gpurun_out
RC_SHARE=1 timeout 90 python profiles/tools/hangcheck.py 3000 1000 > gpurun_out/hangcheck.txt 2>&1
rc=$?; cat gpurun_out/hangcheck.txt | tail -3
if [ $rc -ne 0 ]; then echo "hangcheck rc=$rc: stopping"; exit 1; fi
RC_SHARE=1 timeout 90 python profiles/tools/hangcheck.py 60000 20000 2>&1 | tail -1
run() { timeout 120 python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1  %.1fM configs/s  %.4f ms  e2e %.1fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))" || echo "$1 FAILED"; }
RC_SHARE=0 run share0
RC_SHARE=1 run share1
RC_SHARE=0 run share0
RC_SHARE=1 run share1
RC_SHARE=1 RC_PDL=1 run share1_pdl1
RC_SHARE=0 timeout 120 python profiles/tools/trace.py 100000 508 > gpurun_out/trace_share0.txt 2>&1; tail -14 gpurun_out/trace_share0.txt
RC_SHARE=1 timeout 120 python profiles/tools/trace.py 100000 508 > gpurun_out/trace_share1.txt 2>&1; tail -14 gpurun_out/trace_share1.txt
RC_SHARE=1 timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
