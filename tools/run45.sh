set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu_numa.json 2> gpurun_out/bench_8gpu_numa.err; tail -3 gpurun_out/bench_8gpu_numa.err
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu_numa.json')); print('%.3e'%d['value'], d['ms_per_step'], d['e2e'])"
numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14
