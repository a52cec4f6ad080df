set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_r1_default.err; tail -2 gpurun_out/bench_r1_default.err; wc -l gpurun_out/bench_r1_default.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_default.json')); print('%.3e'%d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_ref.json 2>/dev/null; wc -l gpurun_out/bench_r1_ref.json
