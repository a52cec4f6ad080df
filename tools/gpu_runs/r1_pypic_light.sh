timeout 900 python -m pytest tests/test_gpu_periodic.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -5
python bench.py --workload pypic --steps 40 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('pypic', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], d['config']['picard_iterations_per_step'])"
