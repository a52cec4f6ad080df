timeout 300 python tools/dbg_v6.py 2e7 0 3 2>&1 | grep "^N " | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_sheath.py tests/test_gpu_math.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/dbg_resid.py 4e8 window 8 18 2>&1 | tail -4
timeout 300 python tools/dbg_resid.py 4e8 window 0 3 2>&1 | tail -4
timeout 600 python tools/kbench.py 2e8 window > gpurun_out/kbench4.json 2> gpurun_out/kbench4.err; tail -3 gpurun_out/kbench4.err
