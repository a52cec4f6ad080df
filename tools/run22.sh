timeout 300 python tools/dbg_v6.py 2e7 0 1 2>&1 | grep -v "^   i" | tail -30
