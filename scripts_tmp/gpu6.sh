for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "subdivided_obstacle_matches or global_obstacle" 2>&1 | tail -3; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
