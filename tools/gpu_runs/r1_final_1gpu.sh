set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_r1_default.err; tail -3 gpurun_out/bench_r1_default.err
cat gpurun_out/bench_r1_default.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; cat gpurun_out/bench_r1_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:dd_picard_iter_v6 -s 14 -c 6 -o gpurun_out/prof_r1_v9 $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log | cut -c1-200
for w in explicit pypic boris; do
timeout 600 python bench.py --workload $w --steps 40 --warmup 5 > gpurun_out/bench_r1_$w.json 2> gpurun_out/bench_r1_$w.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_$w.json')); print('$w', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'kernel %.3f ms'%d['roofline']['kernel_ms_mean'], 'share %.2f'%d['roofline']['kernel_share_of_step'])"
done
timeout 600 python tools/bench_paths.py 2e8 gc 4 1 > gpurun_out/paths_gc_final.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open('gpurun_out/paths_gc_final.json'))
for r in d['results']:
    if 'error' in r: print(r); continue
    print("%-55s %9.3f ms  %.3e p-s/s  %7.1f GB/s  frac %.3f" % (r['path'], r['ms'], r['particle_steps_per_s'], r['achieved_gbs'], r['frac']))
PY
