# slab decomposition with the enqueue-ahead Picard loop (gpurun --gpus 2): parity, then slab vs particle decomposition
# at the default size and at a config-5-like size (1e6 cells, 1.25e7 particles per rank)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/slab_check.py 400000 513 > gpurun_out/slab2_ahead.json 2> gpurun_out/slab2_ahead.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/slab2_ahead.err | tail -6; cut -c1-700 gpurun_out/slab2_ahead.json
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$R --master-port 29551 bench.py --gpus 2 --decomposition slab --steps 16 --warmup 3 > gpurun_out/bench_slab2.json 2> gpurun_out/bench_slab2.err
$R --master-port 29552 bench.py --gpus 2 --decomposition slab --steps 16 --warmup 3 --cells 1000000 --total-particles 2.5e7 > gpurun_out/bench_slab2_cfg5.json 2> gpurun_out/bench_slab2_cfg5.err
$R --master-port 29553 bench.py --gpus 2 --steps 16 --warmup 3 --cells 1000000 --total-particles 2.5e7 --no-e2e --no-slab-leg --strong-total 0 > gpurun_out/bench_part2_cfg5.json 2> gpurun_out/bench_part2_cfg5.err
for f in slab2 slab2_cfg5 part2_cfg5; do tail -2 gpurun_out/bench_$f.err | grep -v "^\*\*\*\|OMP_NUM\|^$"; python -c "
import json; d=json.load(open('gpurun_out/bench_$f.json')); print('$f', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.2f'%d['roofline']['kernel_share_of_step'], d['config']['picard_iterations_per_step'], d['config'].get('migration'))"; done
