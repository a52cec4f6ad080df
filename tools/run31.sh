set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 400000 > gpurun_out/mgpu_check2.json 2> gpurun_out/mgpu_check2.err; tail -3 gpurun_out/mgpu_check2.err; cat gpurun_out/mgpu_check2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_2gpu_b.json 2> gpurun_out/bench_2gpu_b.err; tail -3 gpurun_out/bench_2gpu_b.err; cat gpurun_out/bench_2gpu_b.json
