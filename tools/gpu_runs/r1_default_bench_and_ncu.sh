set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_r1_default.err; tail -3 gpurun_out/bench_r1_default.err
cat gpurun_out/bench_r1_default.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; cat gpurun_out/bench_r1_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
ncu --set full --clock-control none --import-source on -k regex:dd_picard_iter_v6 -s 14 -c 6 -o gpurun_out/prof_r1_v8 $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
