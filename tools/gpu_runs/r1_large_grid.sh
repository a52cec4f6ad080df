set -x
timeout 900 python -m pytest tests/test_gpu_sheath.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline --deposit window-big > gpurun_out/bench_big4097.json 2> gpurun_out/bench_big4097.err; tail -3 gpurun_out/bench_big4097.err
timeout 600 python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline --cells 1000000 > gpurun_out/bench_big1e6.json 2> gpurun_out/bench_big1e6.err; tail -3 gpurun_out/bench_big1e6.err
timeout 600 python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline --cells 1000000 --particles-per-gpu 1e8 > gpurun_out/bench_big1e6_1e8.json 2> gpurun_out/bench_big1e6_1e8.err; tail -3 gpurun_out/bench_big1e6_1e8.err
for f in big4097 big1e6 big1e6_1e8; do python -c "
import json; d=json.load(open('gpurun_out/bench_$f.json')); print('$f', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.2f'%d['roofline']['kernel_share_of_step'], d['config']['picard_iterations_per_step'], d['gpu_launches'], d['roofline']['u1_repair_passes'])"; done
