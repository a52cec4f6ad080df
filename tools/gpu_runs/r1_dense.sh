for cells in 4096 512; do for se in 1 8; do
python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline --cells $cells --sort-every $se 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cells', $cells, 'sort_every', $se, '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], d['roofline']['kernel_ms_by_kind'], d['config']['picard_iterations_per_step'])"
done; done
