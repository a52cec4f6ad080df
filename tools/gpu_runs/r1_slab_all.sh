set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/slab_check.py 400000 513 > gpurun_out/slab2.json 2> gpurun_out/slab2.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/slab2.err | tail -12; cat gpurun_out/slab2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/slab_profile.py 2e8 4097 2>/dev/null | grep "^{"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --decomposition slab --steps 16 --warmup 3 > gpurun_out/bench_slab2.json 2> gpurun_out/bench_slab2.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_slab2.err | tail -5
for f in slab2; do python -c "
import json; d=json.load(open('gpurun_out/bench_$f.json')); print('$f', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.2f'%d['roofline']['kernel_share_of_step'], d['config']['picard_iterations_per_step'], d['config']['migration'])"; done
