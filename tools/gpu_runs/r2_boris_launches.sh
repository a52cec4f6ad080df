CMD="python bench.py --workload boris --steps 16 --warmup 3"
$CMD > gpurun_out/plain_boris.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2_boris.csv $CMD > gpurun_out/ncu_l_boris.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2_boris.csv | head -16
