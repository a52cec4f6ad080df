#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29551 tools/slab_phases3.py 1.25e7 1000001 40 2>/dev/null | tee gpurun_out/r2_slab_phases3_cfg5_2gpu.txt
