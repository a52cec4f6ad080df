timeout 900 python -m pytest tests/test_gpu_periodic.py tests/test_gpu_gc.py -m gpu -x -q 2>&1 | tail -40
