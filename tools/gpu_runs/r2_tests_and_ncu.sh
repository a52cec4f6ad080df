# Round 2: the whole GPU suite + smoke, the default bench and the reference arm, then (after each plain run exited 0)
# the ncu launch list of the default step and --set full captures of the dominant kernel and of the lean Boris kernel.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; tail -2 gpurun_out/bench_r2_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; cat gpurun_out/bench_r2_ref.json | cut -c1-600
for wl in boris explicit pypic; do
timeout 600 python bench.py --workload $wl --steps 40 --warmup 3 > gpurun_out/bench_r2_$wl.json 2> gpurun_out/bench_r2_$wl.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$wl.json')); print('$wl', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel ms %.3f'%d['roofline']['kernel_ms_mean'])"
done
timeout 600 python bench.py --workload boris --boris-full-store --steps 40 --warmup 3 > gpurun_out/bench_r2_boris_full.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_boris_full.json')); print('boris full store', '%.3e'%d['value'], 'frac(112B) %.3f'%d['roofline']['frac'], 'kernel ms %.3f'%d['roofline']['kernel_ms_mean'])"
CMD="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-api-leg --strong-total 0"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
CMD2="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-api-leg --strong-total 0"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dd_picard_iter_v6 -s 14 -c 6 -o gpurun_out/prof_r2_v6 $CMD2 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
CMD3="python bench.py --workload boris --steps 2 --warmup 3"
$CMD3 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gc_push_boris_v2 -s 3 -c 2 -o gpurun_out/prof_r2_boris_lean $CMD3 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
