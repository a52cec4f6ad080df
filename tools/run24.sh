set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for se in 4 8; do
timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --sort-every $se > gpurun_out/bench_se$se.json 2> gpurun_out/bench_se$se.err; tail -2 gpurun_out/bench_se$se.err
python -c "
import json; d=json.load(open('gpurun_out/bench_se$se.json')); print('sort_every $se', '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['config']['picard_iterations_per_step'])"
done
CMD="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --sort-every 8"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v6.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
