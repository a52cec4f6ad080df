set -x
timeout 900 python -m pytest tests/test_gpu_sheath.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; tail -3 gpurun_out/bench_e2e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_e2e.json')); print(d['value'], d['e2e'], d['clocks'])"
