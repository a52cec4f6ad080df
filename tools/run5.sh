set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for se in 1 4 8; do
timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --sort-every $se > gpurun_out/bench_se$se.json 2> gpurun_out/bench_se$se.err; tail -3 gpurun_out/bench_se$se.err
python -c "
import json; d=json.load(open('gpurun_out/bench_se$se.json')); print('sort_every $se', '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'])"
done
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --particles-per-gpu 1e8 --sort-every 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dd_picard_iter_v5 -s 6 -c 1 -o gpurun_out/prof_r1_v5 $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
