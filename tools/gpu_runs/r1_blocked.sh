for dep in window window-blocked; do for cells in 4096 512; do
python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline --deposit $dep --cells $cells 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$dep', $cells, '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], d['config']['picard_iterations_per_step'])"
done; done
