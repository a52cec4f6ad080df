#!/bin/bash
# distributed field update of the slab decomposition on 2 GPUs: parity (tools/slab_check.py, both field updates), then
# BASELINE config 5's shape (1e6 cells, 1.25e7 particles per GPU) slab distributed / slab replicated / particle decomposition
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "slab" 2>&1 | tail -3
for f in distributed replicated; do
PIC_SLAB_FIELD=$f $TR --master-port 29541 tools/slab_check.py 400000 513 2>gpurun_out/slab_check2_$f.err | grep '^{' | tee gpurun_out/r2_slab_check2_$f.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('slab_check $f', {k: d[k] for k in ('ok','iters_slab','iters_single','E_rel','electrons_x_rel','ions_x_rel','stat')})"
done
one() { python -c "
import json,sys; d=json.load(sys.stdin); print('$1:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'k', d['config'].get('picard_iterations_per_step', d['roofline'].get('mean_picard_iterations')), 'kernel ms %.4f'%d['roofline']['kernel_ms_mean'])"; }
for rep in 1 2; do
for f in distributed replicated; do
$TR --master-port 29542 bench.py --gpus 2 --decomposition slab --slab-field $f --cells 1000000 --particles-per-gpu 1.25e7 --steps 16 --warmup 3 --sort-every 8 2>gpurun_out/bench_slab2_$f.err | tee gpurun_out/bench_slab2_cfg5_$f.json | one "cfg5 x2 slab $f"
done
$TR --master-port 29543 bench.py --gpus 2 --cells 1000000 --particles-per-gpu 1.25e7 --steps 16 --warmup 3 --sort-every 8 --no-e2e --no-cpu-baseline --strong-total 0 --no-slab-leg --no-api-leg 2>gpurun_out/bench_part2.err | tee gpurun_out/bench_part2_cfg5.json | one "cfg5 x2 particle decomposition"
done
# the bench workload (4097 nodes, 2e8 particles per GPU) on slabs, both field updates
for f in distributed replicated; do
$TR --master-port 29544 bench.py --gpus 2 --decomposition slab --slab-field $f --steps 16 --warmup 3 --sort-every 8 2>gpurun_out/bench_slab2b_$f.err | tee gpurun_out/bench_slab2_default_$f.json | one "4097 nodes x2 slab $f"
done
