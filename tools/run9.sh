set -x
timeout 900 python -m pytest tests/test_gpu_sheath.py tests/test_gpu_math.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --sort-every 8 > gpurun_out/bench_v5b.json 2> gpurun_out/bench_v5b.err; tail -3 gpurun_out/bench_v5b.err
cat gpurun_out/bench_v5b.json
