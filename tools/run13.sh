timeout 300 python tools/dbg_sorted.py 2>&1 | tail -30
