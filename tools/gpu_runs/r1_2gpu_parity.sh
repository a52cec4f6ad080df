set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 400000 > gpurun_out/mgpu_check2.json 2> gpurun_out/mgpu_check2.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/mgpu_check2.err | tail -8; cat gpurun_out/mgpu_check2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload boris --steps 40 --warmup 5 > gpurun_out/bench_boris2.json 2> gpurun_out/bench_boris2.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_boris2.err | tail -5
python bench.py --workload boris --steps 40 --warmup 5 > gpurun_out/bench_r1_boris.json 2>/dev/null
for f in bench_boris2 bench_r1_boris; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.2f'%d['roofline']['kernel_share_of_step'])"; done
