timeout 300 python scripts_tmp/feed_big.py 2>&1 | tail -4
