set -x
python tools/slab_check.py 400000 513 > gpurun_out/slab1.json 2> gpurun_out/slab1.err; tail -5 gpurun_out/slab1.err; cat gpurun_out/slab1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/slab_check.py 400000 513 > gpurun_out/slab2.json 2> gpurun_out/slab2.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/slab2.err | tail -12; cat gpurun_out/slab2.json
