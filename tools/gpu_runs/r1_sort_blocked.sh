# blocked + unrolled global-cursor sort: timing, then the tests that cover it, then the bench
python tools/sortbench.py 2e8 4097 32 2>gpurun_out/sortb32.err | tee gpurun_out/sortb32.json
python tools/sortbench.py 2e8 4097 0 2>gpurun_out/sortb0.err | tee gpurun_out/sortb0.json
timeout 600 python -m pytest tests/test_gpu_sheath.py tests/test_gpu_edge_cases.py tests/test_gpu_periodic.py tests/test_gpu_gc.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_sortb.err | tee gpurun_out/bench_sortb.json | python -c "
import json,sys; d=json.load(sys.stdin); print('bench', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share', '%.3f'%d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz'])"
