set -x
for w in explicit pypic boris; do
timeout 600 python bench.py --workload $w --steps 40 --warmup 3 > gpurun_out/bench_r1_$w.json 2> gpurun_out/bench_r1_$w.err; tail -3 gpurun_out/bench_r1_$w.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_mean'], d['roofline']['kernel_share_of_step'], d['config']['picard_iterations_per_step'], d['clocks'])"
done
