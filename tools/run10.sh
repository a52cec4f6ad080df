set -x
timeout 900 python -m pytest tests/test_gpu_sheath.py tests/test_gpu_math.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/kbench.py 2e8 window,window-ldg > gpurun_out/kbench1.json 2> gpurun_out/kbench1.err; tail -5 gpurun_out/kbench1.err
timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --sort-every 8 > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; tail -3 gpurun_out/bench_v6.err
cat gpurun_out/bench_v6.json
