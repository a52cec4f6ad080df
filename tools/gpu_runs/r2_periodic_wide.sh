#!/bin/bash
# 15-node deposit windows in the periodic Picard and the explicit window kernels (chosen at run time; PIC_S_NARROW=1
# keeps the 7-node build): full GPU suite, then A/B of both workloads over the sort interval
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() {  # label env workload sort_every
  env $2 python bench.py --workload $3 --steps 48 --warmup 4 --sort-every $4 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('$3 $1 sort every $4:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'kernel ms %.4f'%r['kernel_ms_mean'], 'frac %.3f'%r['frac'], d['clocks']['sm_mhz'])"
}
for rep in 1 2; do
for wl in explicit pypic; do
run "narrow" PIC_S_NARROW=1 $wl 8
run "wide" PIC_S_NARROW=0 $wl 8
run "wide" PIC_S_NARROW=0 $wl 12
run "wide" PIC_S_NARROW=0 $wl 16
done
done 2>&1 | tee gpurun_out/r2_periodic_wide.txt
