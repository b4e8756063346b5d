timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "million" 2>&1 | tail -1
