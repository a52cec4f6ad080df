timeout 900 python -m pytest tests/test_gpu_sheath.py tests/test_gpu_dropin.py tests/test_gpu_edge_cases.py tests/test_checkpoint.py -m gpu -x -q 2>&1 | tail -5
for rep in 1 2; do
python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('sheath', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], d['roofline']['kernel_ms_by_kind'], 'repairs', d['roofline']['u1_repair_passes'], d['clocks']['sm_mhz'])"
done
