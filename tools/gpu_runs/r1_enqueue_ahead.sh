# enqueue-ahead Picard loop: sheath tests (goldens included), then the bench A/B is the previous commit's number
timeout 900 python -m pytest tests/test_gpu_sheath.py tests/test_gpu_reproducible.py tests/test_gpu_dropin.py tests/test_gpu_edge_cases.py tests/test_checkpoint.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8
for rep in 1 2; do
python bench.py --steps 60 --warmup 4 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_ahead.err | tee gpurun_out/bench_ahead.json | python -c "
import json,sys; d=json.load(sys.stdin); print('bench', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share', '%.3f'%d['roofline']['kernel_share_of_step'], d['roofline']['kernel_ms_by_kind'], d['roofline']['u1_repair_passes'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done
