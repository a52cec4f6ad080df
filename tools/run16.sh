set -x
timeout 900 python bench.py --steps 16 --warmup 3 > gpurun_out/bench_v6c.json 2> gpurun_out/bench_v6c.err; tail -3 gpurun_out/bench_v6c.err
cat gpurun_out/bench_v6c.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --particles-per-gpu 1e8 --sort-every 8"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v6.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
