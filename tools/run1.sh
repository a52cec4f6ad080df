set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
timeout 300 python __graft_entry__.py 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --particles-per-gpu 2e7 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
