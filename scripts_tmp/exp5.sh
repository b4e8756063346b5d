#!/bin/bash
python profiles/tools/small.py final
RC_SHARE=1 RC_SPLIT_MAX=16 python profiles/tools/small.py split16_share
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "planner_sized or scheduling_variants or single_launch or ragged or culled_equals_brute or gather_two_ranks or different_streams or million" 2>&1 | tail -2
