python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/debug/profile_tracked.py 2e7 100 nosync 2>&1 | head -1
python tools/debug/profile_tracked.py 2e7 100 sync 2>&1 | head -1
python tools/debug/profile_tracked.py 2e8 60 nosync 2>&1 | head -1
