set -x
timeout 900 python -m pytest tests/test_gpu_init.py tests/test_checkpoint.py -m gpu -x -q 2>&1 | tail -15
