for rep in 1 2; do for h in 1 8; do
python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu-baseline --heavy-sort-every $h 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('sheath heavy_sort_every', $h, '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.3f'%d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz'])"
done; done
