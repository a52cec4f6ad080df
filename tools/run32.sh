set -x
timeout 600 python tools/sortbench.py 2e8
timeout 900 python -m pytest tests/test_gpu_sheath.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 40 --no-cpu-baseline --no-e2e > gpurun_out/bench_sort2.json 2> gpurun_out/bench_sort2.err; tail -3 gpurun_out/bench_sort2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_sort2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'])"
