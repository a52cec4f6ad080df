#!/bin/bash
# mixed-species fused Boris kernel: parity tests, then timing against the v1 kernels
python -m pytest tests/test_gpu_gc.py -x -q -m gpu 2>&1 | tail -15
python tools/debug/profile_boris_mixed.py 1e8 16
