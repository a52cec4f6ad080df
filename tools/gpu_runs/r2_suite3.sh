#!/bin/bash
# full GPU suite after the tail folding / merged prologue, then every workload's bench line
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for wl in explicit pypic boris; do
python bench.py --workload $wl --steps 40 --warmup 3 2>/dev/null > gpurun_out/bench_r2b_$wl.json
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r2b_$wl.json')); print('$wl', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel ms %.3f'%d['roofline']['kernel_ms_mean'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
python bench.py --steps 60 --warmup 3 --no-e2e --no-cpu-baseline --strong-total 0 2>/dev/null > gpurun_out/bench_r2b_sheath.json
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r2b_sheath.json')); print('sheath', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'share %.3f'%d['roofline']['kernel_share_of_step'], 'api %.3e %.3f ms'%(d['reference_api']['value'], d['reference_api']['ms_per_step']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
