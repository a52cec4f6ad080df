python -m pytest tests/test_gpu_gc.py tests/test_gpu_dropin.py tests/test_gpu_init.py tests/test_gpu_reproducible.py tests/test_gpu_periodic.py -m gpu -x -q 2>&1 | tail -5
for extra in "" "--boris-full-store"; do
python bench.py --workload boris --steps 40 --warmup 3 $extra 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('boris $extra', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel ms %.3f'%d['roofline']['kernel_ms_mean'], d['clocks']['sm_mhz'])"
done
