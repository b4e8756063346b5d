#!/bin/bash
export RC_LIB=$PWD/ab/librobocheck_w28.so
for W in 32 28 32 28; do RC_CULLED_WARPS=$W timeout 200 python profiles/tools/quick.py w$W 2>&1 | grep -E "hard100k|hard1M|hard10k\"" | cut -c1-140; done
RC_CULLED_WARPS=28 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "culled_matches or culled_equals_brute or million or full_size_workloads" 2>&1 | tail -1
