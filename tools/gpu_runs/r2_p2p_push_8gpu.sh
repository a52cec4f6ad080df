# push-based peer-memory reduction on 8 GPUs: parity against the NCCL path, then A/B of the weak-scaling line
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
  tools/p2p_check.py 1600000 6 > gpurun_out/p2p_push_check8.json 2> gpurun_out/p2p_push_check8.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/p2p_push_check8.json')); print('ok', d['ok'], [r['E_rel'] for r in d['ranks']][:2], d['ranks'][0]['nccl_vs_nccl'])"; tail -c 400 gpurun_out/p2p_push_check8.err
for red in nccl p2p nccl p2p; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 \
  bench.py --gpus 8 --steps 60 --warmup 3 --no-e2e --no-cpu-baseline --no-api-leg --no-slab-leg --strong-total 0 --reduce $red 2>gpurun_out/bench_8gpu_push_$red.err | tee gpurun_out/bench_8gpu_push_$red.json | python -c "
import json,sys; d=json.load(sys.stdin); print('$red', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'share %.3f'%d['roofline']['kernel_share_of_step'], 'kernel %.4f'%d['roofline']['kernel_ms_mean'], d['clocks']['sm_mhz'])"
done
