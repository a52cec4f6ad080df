#!/bin/bash
# after the reproducible builds of the periodic kernels: full GPU suite, the default builds' speed (explicit, pypic, sheath)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for wl in explicit pypic; do
python bench.py --workload $wl --steps 48 --warmup 4 2>/dev/null > gpurun_out/bench_r2_final2_$wl.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_final2_$wl.json')); print('$wl %.3e %.3f ms/step kernel ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac']), d['clocks']['sm_mhz'])"
done
python bench.py --steps 48 --warmup 3 --no-e2e --no-cpu-baseline --strong-total 0 2>/dev/null > gpurun_out/bench_r2_final2_sheath.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_final2_sheath.json')); r=d['roofline']; print('sheath %.3e %.3f ms/step frac %.3f share %.3f api %.3e' % (d['value'], d['ms_per_step'], r['frac'], r['kernel_share_of_step'], d['reference_api']['value']), d['clocks']['sm_mhz'])"
