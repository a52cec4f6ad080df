set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err; tail -3 gpurun_out/bench_warp.err
python -c "
import json; d=json.load(open('gpurun_out/bench_warp.json')); print('warp', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'])"
