# reproducible build (window-det): parity tests, then A/B against the default build on one box
timeout 600 python -m pytest tests/test_gpu_reproducible.py -m gpu -x -q 2>&1 | tail -15
for dep in window window-det; do
python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu-baseline --deposit $dep 2>gpurun_out/bench_$dep.err | tee gpurun_out/bench_$dep.json | python -c "
import json,sys; d=json.load(sys.stdin); print('$dep', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], d['roofline']['kernel_ms_by_kind'], 'share', '%.3f'%d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz'])"
done
