set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v6d.json 2> gpurun_out/bench_v6d.err; tail -3 gpurun_out/bench_v6d.err
cat gpurun_out/bench_v6d.json
