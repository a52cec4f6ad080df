set -x
nvidia-smi -L | head -8; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py 800000 > gpurun_out/mgpu_check8.json 2> gpurun_out/mgpu_check8.err; tail -2 gpurun_out/mgpu_check8.err; cat gpurun_out/mgpu_check8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -3 gpurun_out/bench_8gpu.err; cat gpurun_out/bench_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --total-particles 1e9 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_8gpu_1e9.json 2> gpurun_out/bench_8gpu_1e9.err; tail -3 gpurun_out/bench_8gpu_1e9.err; cat gpurun_out/bench_8gpu_1e9.json
