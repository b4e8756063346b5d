python scripts_tmp/stress.py 2>&1 | tail -20
