set -x
CMD="python bench.py --workload boris --steps 4 --warmup 3"
$CMD > gpurun_out/plain_boris.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_boris.csv $CMD > gpurun_out/ncu_lb.log 2>&1
tail -2 gpurun_out/ncu_lb.log
CMD="python bench.py --workload explicit --steps 4 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_explicit.csv $CMD > gpurun_out/ncu_le.log 2>&1
tail -2 gpurun_out/ncu_le.log
