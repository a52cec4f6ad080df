set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
for v in warp atomic; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --deposit $v > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; tail -3 gpurun_out/bench_$v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$v.json')); print('$v', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'])"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --sort-every 1 > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_s1.json')); print('sort1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --sort-every 0 > gpurun_out/bench_s0.json 2> gpurun_out/bench_s0.err; python -c "
import json; d=json.load(open('gpurun_out/bench_s0.json')); print('nosort', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'])"
