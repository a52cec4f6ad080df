# sort interval A/B with the faster re-sort (1.9 ms): 6 vs 8 steps, same box, interleaved
for se in 6 8 6 8 5; do
python bench.py --sort-every $se --steps 48 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('sort_every=$se', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.4f'%d['roofline']['kernel_ms_mean'], 'share %.3f'%d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz'])"
done
