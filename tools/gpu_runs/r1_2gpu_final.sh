set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_2gpu.err | tail -3
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('2gpu', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['u1_repair_passes'])"
