# pypic enqueue-ahead loop + canary tests, then the pypic workload bench
timeout 600 python -m pytest tests/test_gpu_periodic.py tests/test_gpu_canaries.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -6
python bench.py --workload pypic --steps 40 --warmup 4 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_pypic_ahead.err | tee gpurun_out/bench_pypic_ahead.json | python -c "
import json,sys; d=json.load(sys.stdin); print('pypic', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], d['roofline'].get('kernel_ms_mean'), d['roofline'].get('kernel_share_of_step'), d['clocks']['sm_mhz'])"
