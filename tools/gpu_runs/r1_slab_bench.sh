set -x
timeout 600 python bench.py --decomposition slab --steps 16 --warmup 3 > gpurun_out/bench_slab1.json 2> gpurun_out/bench_slab1.err; tail -3 gpurun_out/bench_slab1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --decomposition slab --steps 16 --warmup 3 > gpurun_out/bench_slab2.json 2> gpurun_out/bench_slab2.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_slab2.err | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --decomposition slab --steps 16 --warmup 3 --cells 1000000 --total-particles 1e8 > gpurun_out/bench_slab2_1e6.json 2> gpurun_out/bench_slab2_1e6.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_slab2_1e6.err | tail -5
for f in slab1 slab2 slab2_1e6; do python -c "
import json; d=json.load(open('gpurun_out/bench_$f.json')); print('$f', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.2f'%d['roofline']['kernel_share_of_step'], d['config']['picard_iterations_per_step'], d['config']['migration'])"; done
