timeout 500 python tools/dbg_race.py 40 2>&1 | tail -60
