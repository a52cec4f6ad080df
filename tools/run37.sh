set -x
timeout 900 python -m pytest tests/test_gpu_gc.py tests/test_gpu_sheath.py -m gpu -x -q 2>&1 | tail -5
for w in boris; do
timeout 600 python bench.py --workload $w --steps 40 --warmup 5 > gpurun_out/bench_r1_$w.json 2> gpurun_out/bench_r1_$w.err; tail -3 gpurun_out/bench_r1_$w.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_$w.json')); print('$w', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.2f'%d['roofline']['kernel_share_of_step'], d['config']['picard_iterations_per_step'], d['gpu_launches'])"
done
