set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; tail -3 gpurun_out/bench_v7.err
cat gpurun_out/bench_v7.json
