timeout 1200 python scripts_tmp/c5knobs.py 2>&1 | grep -v "^$"
