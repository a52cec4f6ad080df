timeout 1200 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_gc.py tests/test_gpu_periodic.py -m gpu -q 2>&1 | tail -40
