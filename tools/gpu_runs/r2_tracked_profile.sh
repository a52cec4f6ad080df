python -m pytest tests/test_gpu_sheath.py tests/test_gpu_dropin.py tests/test_checkpoint.py -m gpu -x -q 2>&1 | tail -6
python tools/debug/profile_tracked.py 2e7 100 sync 2>&1 | head -60
python tools/debug/profile_tracked.py 2e7 100 nosync 2>&1 | head -2
python tools/debug/profile_tracked.py 2e8 60 nosync 2>&1 | head -2
python tools/debug/profile_tracked.py 2e8 60 sync 2>&1 | head -2
