#!/bin/bash
# default builds of the periodic kernels after the exact routines went back to their original signatures
python -m pytest tests/test_gpu_reproducible.py tests/test_gpu_periodic.py -x -q -m gpu 2>&1 | tail -3
for wl in explicit pypic; do
python bench.py --workload $wl --steps 48 --warmup 4 2>/dev/null > gpurun_out/bench_r2_final3_$wl.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_final3_$wl.json')); print('$wl %.3e %.3f ms/step kernel ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac']), d['clocks']['sm_mhz'])"
done
