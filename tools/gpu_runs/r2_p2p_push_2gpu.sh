# push-based peer-memory reduction (gpurun --gpus 2): parity against the NCCL path, the enabled test, A/B of the weak-scaling line
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
  tools/p2p_check.py 400000 6 > gpurun_out/p2p_push_check2.json 2> gpurun_out/p2p_push_check2.err; echo rc=$?; cat gpurun_out/p2p_push_check2.json; tail -c 800 gpurun_out/p2p_push_check2.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
for red in nccl p2p nccl p2p; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 \
  bench.py --gpus 2 --steps 60 --warmup 3 --no-e2e --no-cpu-baseline --no-api-leg --no-slab-leg --strong-total 0 --reduce $red 2>gpurun_out/bench_2gpu_push_$red.err | tee gpurun_out/bench_2gpu_push_$red.json | python -c "
import json,sys; d=json.load(sys.stdin); print('$red', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'share %.3f'%d['roofline']['kernel_share_of_step'], 'kernel %.4f'%d['roofline']['kernel_ms_mean'], d['clocks']['sm_mhz'])"
done
