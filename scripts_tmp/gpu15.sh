python scripts_tmp/c5knobs.py 2>&1 | tail -8
