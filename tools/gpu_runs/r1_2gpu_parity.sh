set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 400000 > gpurun_out/mgpu_check2.json 2> gpurun_out/mgpu_check2.err; tail -5 gpurun_out/mgpu_check2.err; cat gpurun_out/mgpu_check2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/b2.json 2>/dev/null; wc -l gpurun_out/b2.json
