# Round 2, one 8-GPU box: peer-memory reduction at 8 ranks (parity + A/B), the strong-scaling sub-record, slabs.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29571 tools/p2p_check.py 1600000 6 > gpurun_out/p2p_check8.json 2> gpurun_out/p2p_check8.err; echo p2p_check rc=$?; cat gpurun_out/p2p_check8.json | cut -c1-900
for red in nccl p2p; do
timeout 400 $TR --nproc-per-node 8 --master-port 29572 bench.py --gpus 8 --steps 60 --warmup 3 --no-e2e --reduce $red 2>gpurun_out/bench_8gpu_$red.err > gpurun_out/bench_8gpu_$red.json
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu_$red.json')); s=d['strong_scaling']; print('$red weak %.3e %.3f ms share %.3f | strong 1e9: %.3e %.3f ms share %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_share_of_step'], s['value'], s['ms_per_step'], s['kernel_share_of_step']))"
tail -2 gpurun_out/bench_8gpu_$red.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29573 tools/slab_check.py 1600000 4097 > gpurun_out/slab_check8.json 2> gpurun_out/slab_check8.err; echo slab rc=$?; tail -c 600 gpurun_out/slab_check8.json
timeout 400 $TR --nproc-per-node 8 --master-port 29574 bench.py --gpus 8 --steps 20 --warmup 3 --reduce nccl 2>gpurun_out/bench_8gpu_full.err > gpurun_out/bench_8gpu_full.json
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu_full.json')); print('full line e2e', d['e2e'])"
