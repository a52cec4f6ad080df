# 2 GPUs: sharded parity (sheath incl. reproducible build, pypic, explicit, Boris), then the weak-scaling bench line
timeout 280 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "two_rank_sheath" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('2gpu', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'share %.3f'%d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz'], d.get('e2e',{}) and '%.3e'%d['e2e']['value'])"
