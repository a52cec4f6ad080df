# final single-GPU evidence of the round: full GPU suite, smoke, default bench, launch list, long run, reproducible build
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_r1_default.err; tail -2 gpurun_out/bench_r1_default.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_default.json')); print('default', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'share %.3f'%d['roofline']['kernel_share_of_step'], 'e2e %.3e'%d['e2e']['value'], 'cpu %.3e'%d['cpu_baseline']['value'], d['clocks'])"
CMD="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_final.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log
python bench.py --steps 1000 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_1000steps.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_1000steps.json')); print('1000 steps', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], d['clocks'])"
python bench.py --deposit window-det --steps 100 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_window_det.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_window_det.json')); print('det', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], d['clocks']['sm_mhz'])"
