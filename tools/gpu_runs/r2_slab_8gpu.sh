#!/bin/bash
# BASELINE config 5 on one 8-GPU box: 1e6-cell grid, 1e8 particles, 8 subdomains.  Slab decomposition with the
# distributed field update vs the replicated one vs the particle decomposition; parity of the slabs first.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29573 tools/slab_check.py 1600000 4097 > gpurun_out/r2_slab_check8_distributed.json 2> gpurun_out/slab_check8.err; echo slab_check rc=$?; tail -c 700 gpurun_out/r2_slab_check8_distributed.json
one() { python -c "
import json,sys; d=json.load(sys.stdin); print('$1:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'k', d['config'].get('picard_iterations_per_step', d['roofline'].get('mean_picard_iterations')), 'kernel ms %.4f'%d['roofline']['kernel_ms_mean'])"; }
for rep in 1 2; do
timeout 200 $TR --master-port 29542 bench.py --gpus 8 --decomposition slab --cells 1000000 --total-particles 1e8 --steps 24 --warmup 4 --sort-every 8 2>gpurun_out/err_slab8.txt | tee gpurun_out/r2_bench_cfg5_8gpu_slab.json | one "cfg5 x8 slab distributed"
timeout 200 $TR --master-port 29543 bench.py --gpus 8 --cells 1000000 --total-particles 1e8 --steps 24 --warmup 4 --sort-every 8 --no-e2e --no-cpu-baseline --strong-total 0 --no-slab-leg --no-api-leg 2>gpurun_out/err_part8.txt | tee gpurun_out/r2_bench_cfg5_8gpu_particle.json | one "cfg5 x8 particle decomposition"
done
timeout 200 $TR --master-port 29544 bench.py --gpus 8 --decomposition slab --slab-field replicated --cells 1000000 --total-particles 1e8 --steps 24 --warmup 4 --sort-every 8 2>/dev/null | tee gpurun_out/r2_bench_cfg5_8gpu_slab_replicated.json | one "cfg5 x8 slab replicated"
timeout 100 $TR --master-port 29551 tools/slab_phases.py 1.25e7 1000001 8 2>/dev/null | grep '^{' | head -2 | tee gpurun_out/r2_slab_phases_cfg5_8gpu.txt | cut -c1-600
# the bench workload (4097 nodes, 2e8 particles per GPU) on 8 slabs
timeout 200 $TR --master-port 29545 bench.py --gpus 8 --decomposition slab --steps 16 --warmup 4 2>/dev/null | tee gpurun_out/r2_bench_default_8gpu_slab.json | one "4097 nodes x8 slab distributed"
