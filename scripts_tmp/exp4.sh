#!/bin/bash
P="python profiles/tools/small.py"
export RC_LIB=$PWD/ab/librobocheck_v3.so
RC_SPLIT_MAX=16 $P split16
RC_SPLIT_MAX=32 $P split32
RC_SPLIT_MAX=64 $P split64
RC_SPLIT_MAX=64 RC_SPLIT_LOAD=16 $P split64_load16
RC_SPLIT_MAX=64 RC_SHARE=0 $P split64_noshare
RC_SPLIT_MAX=16 $P split16
RC_SPLIT_MAX=64 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "planner_sized or scheduling_variants or single_launch" 2>&1 | tail -2
