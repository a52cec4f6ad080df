#!/bin/bash
python -m pytest tests/test_gpu_sheath.py tests/test_gpu_dropin.py tests/test_gpu_reproducible.py tests/test_gpu_init.py tests/test_checkpoint.py -x -q -m gpu 2>&1 | tail -8
python tools/debug/ab_prologue.py 2e8 40
python tools/debug/ab_prologue.py 2e7 80
