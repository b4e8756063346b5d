#!/bin/bash
for i in 1 2 3 4 5; do
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_launch or planner_sized or scheduling_variants or gather_two_ranks or different_streams or pipelined or check_host or subdivided_obstacle_matches or stack_overflow" 2>&1 | tail -1
done
