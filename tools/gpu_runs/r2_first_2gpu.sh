# First GPU call of round 2 (gpurun --gpus 2): the peer-memory reduction written at the end of round 1.
#  1. parity against the NCCL path without sorting (bit-identical fields expected at 2 ranks)
#  2. A/B of the weak-scaling line: NCCL all-reduce vs reduction inside the field kernel
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
  tools/p2p_check.py 400000 6 > gpurun_out/p2p_check2.json 2> gpurun_out/p2p_check2.err; echo rc=$?; cat gpurun_out/p2p_check2.json; tail -c 800 gpurun_out/p2p_check2.err
PIC_TEST_P2P=1 timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
for red in nccl p2p nccl p2p; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 \
  bench.py --gpus 2 --steps 60 --warmup 3 --no-e2e --no-cpu-baseline --reduce $red 2>gpurun_out/bench_2gpu_$red.err | tee gpurun_out/bench_2gpu_$red.json | python -c "
import json,sys; d=json.load(sys.stdin); print('$red', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'share %.3f'%d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz'])"
done
