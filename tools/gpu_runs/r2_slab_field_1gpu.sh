#!/bin/bash
# distributed field update of the slab decomposition: emulated-rank parity on one GPU + the drop-in loops on the sorted store
python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -x -q -m gpu -k "slab or pic_l_main_module or pypic_main_module" 2>&1 | tail -15
# config 5's per-rank share on one GPU, both field updates (world 1: no collectives; the field kernels' cost)
for f in distributed replicated; do
python bench.py --decomposition slab --slab-field $f --cells 1000000 --particles-per-gpu 1.25e7 --steps 12 --warmup 3 --sort-every 8 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('slab $f 1 GPU 1e6 cells 1.25e7 particles:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'k', d['config']['picard_iterations_per_step'], 'kernel ms %.4f'%d['roofline']['kernel_ms_mean'])"
done
