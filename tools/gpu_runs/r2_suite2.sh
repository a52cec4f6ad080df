python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --workload boris --steps 40 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('boris lean', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel ms %.3f'%d['roofline']['kernel_ms_mean'], d['clocks']['sm_mhz'])"
python bench.py --steps 60 --warmup 3 --no-e2e --no-cpu-baseline --strong-total 0 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('sheath', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'share %.3f'%d['roofline']['kernel_share_of_step'], 'api %.3e %.3f ms'%(d['reference_api']['value'], d['reference_api']['ms_per_step']), d['clocks']['sm_mhz'])"
python tools/debug/profile_tracked.py 2e7 100 nosync 2>&1 | head -1
