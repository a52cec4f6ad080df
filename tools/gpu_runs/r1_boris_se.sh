timeout 600 python -m pytest tests/test_gpu_gc.py -m gpu -x -q 2>&1 | tail -3
for se in 6 8 10 12; do
python bench.py --workload boris --steps 48 --warmup 5 --sort-every $se 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('boris sort interval', $se//2, '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'frac %.3f'%d['roofline']['frac'])"
done
