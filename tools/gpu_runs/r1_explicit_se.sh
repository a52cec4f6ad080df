timeout 600 python -m pytest tests/test_gpu_periodic.py -m gpu -x -q 2>&1 | tail -3
for w in explicit pypic; do for se in 8 12 16; do
python bench.py --workload $w --steps 48 --warmup 3 --sort-every $se 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$w sort_every', $se, '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'frac %.3f'%d['roofline']['frac'], d['config']['picard_iterations_per_step'])"
done; done
