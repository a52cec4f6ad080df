#!/bin/bash
# replicated re-injection draws of the slab decomposition: parity on 2 GPUs, then the step's sections
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "slab_sim_single_rank" 2>&1 | tail -2
$TR --master-port 29541 tools/slab_check.py 400000 513 2>gpurun_out/slab_check2.err | grep '^{' | tee gpurun_out/r2_slab_check2_replicated_draws.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('slab_check', {k: d[k] for k in ('ok','iters_slab','iters_single','E_rel','electrons_x_rel','ions_x_rel','electrons_dead','ions_dead','stat','total_particles')})"
tail -3 gpurun_out/slab_check2.err
$TR --master-port 29551 tools/slab_phases.py 1.25e7 1000001 16 2>/dev/null | grep '^{' | head -1 | cut -c1-700 | tee gpurun_out/r2_slab_phases_cfg5_2gpu_v2.txt
$TR --master-port 29552 tools/slab_phases.py 2e8 4097 16 2>/dev/null | grep '^{' | head -1 | cut -c1-700 | tee gpurun_out/r2_slab_phases_default_2gpu_v2.txt
