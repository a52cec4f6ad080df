#!/bin/bash
# config 3 full-size Boris parity test; launch list of a slab step (world 1, config 5's per-rank share)
python -m pytest tests/test_gpu_gc.py -x -q -m gpu -k "baseline_config_3" 2>&1 | tail -12
CMD="python bench.py --decomposition slab --cells 1000000 --particles-per-gpu 1.25e7 --steps 3 --warmup 2 --sort-every 8"
$CMD > gpurun_out/plain_slab.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2_slab_w1.csv $CMD > gpurun_out/ncu_l_slab.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2_slab_w1.csv | head -24 | tee gpurun_out/r2_launches_slab_w1.txt
