timeout 300 python tools/dbg_resid.py 4e8 window 8 20 2>&1 | tail -12
timeout 300 python tools/dbg_resid.py 4e8 window-ldg 8 12 2>&1 | tail -8
timeout 300 python tools/dbg_resid.py 4e8 warp 8 12 2>&1 | tail -8
timeout 300 python tools/dbg_resid.py 4e8 window 0 6 2>&1 | tail -8
