#!/bin/bash
# slab (distributed field update) vs particle decomposition on 2 GPUs over the grid size at ~25 particles per cell,
# and the bench workload (4097 nodes, 2e8 particles per GPU)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "slab" 2>&1 | tail -2
one() { python -c "
import json,sys; d=json.load(sys.stdin); print('$1:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'k', d['config'].get('picard_iterations_per_step', d['roofline'].get('mean_picard_iterations')), 'kernel ms %.4f'%d['roofline']['kernel_ms_mean'])"; }
for cfg in "1000000 1.25e7" "4000000 5e7" "16000000 2e8"; do
set -- $cfg
$TR --master-port 29542 bench.py --gpus 2 --decomposition slab --cells $1 --particles-per-gpu $2 --steps 8 --warmup 3 --sort-every 8 2>gpurun_out/err_slab.txt | tee gpurun_out/bench_x2_slab_$1.json | one "$1 cells x2 slab"
$TR --master-port 29543 bench.py --gpus 2 --cells $1 --particles-per-gpu $2 --steps 8 --warmup 3 --sort-every 8 --no-e2e --no-cpu-baseline --strong-total 0 --no-slab-leg --no-api-leg 2>gpurun_out/err_part.txt | tee gpurun_out/bench_x2_part_$1.json | one "$1 cells x2 particle decomposition"
done
for f in distributed replicated; do
$TR --master-port 29544 bench.py --gpus 2 --decomposition slab --slab-field $f --steps 16 --warmup 3 --sort-every 8 2>gpurun_out/bench_slab2b_$f.err | tee gpurun_out/bench_slab2_default_$f.json | one "4097 nodes x2 slab $f"
done
tail -3 gpurun_out/err_slab.txt gpurun_out/err_part.txt
