set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 24 --warmup 3 --no-e2e > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_8gpu.err | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --decomposition slab --steps 24 --warmup 3 > gpurun_out/bench_slab8.json 2> gpurun_out/bench_slab8.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_slab8.err | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/slab_profile.py 2e8 4097 2>/dev/null | grep "^{" | sort | head -8
for f in 8gpu slab8; do python -c "
import json; d=json.load(open('gpurun_out/bench_$f.json')); print('$f', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.2f'%d['roofline']['kernel_share_of_step'], d['config']['picard_iterations_per_step'], d['config'].get('kernel_ms_and_particles_by_rank'))"; done
