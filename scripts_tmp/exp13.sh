#!/bin/bash
python profiles/tools/small.py auto
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "planner_sized or scheduling_variants or single_launch or ragged or culled_equals_brute or gather_two_ranks or different_streams or million or culled_matches or golden or full_size" 2>&1 | tail -2
echo "== C5 with fewer warps"
for W in 32 24 16; do RC_CULLED_WARPS=$W timeout 300 python profiles/tools/quick.py w$W --c5 2>&1 | grep -E "c5" | cut -c1-150; done
