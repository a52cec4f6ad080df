set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for se in 1 4 8 16; do
timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --sort-every $se > gpurun_out/bench_se$se.json 2> gpurun_out/bench_se$se.err; tail -3 gpurun_out/bench_se$se.err
python -c "
import json; d=json.load(open('gpurun_out/bench_se$se.json')); print('sort_every $se', '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'])"
done
