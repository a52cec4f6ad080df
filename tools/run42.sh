set -x
timeout 900 python -m pytest tests/test_gpu_sheath.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/sortbench.py 2e8 1000001
