for n in 5e7 1e8 2e8 4e8; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --particles-per-gpu $n > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -2 gpurun_out/bench_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print('N $n', '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'], d['config']['picard_iterations_per_step'])"
done
