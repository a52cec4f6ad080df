set -x
timeout 900 python -m pytest tests/test_gpu_sheath.py tests/test_gpu_math.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/kbench.py 2e8 window > gpurun_out/kbench2.json 2> gpurun_out/kbench2.err; tail -5 gpurun_out/kbench2.err
