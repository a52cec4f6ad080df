# launch lists of the other movers' full steps (where does the time outside the particle kernel go?)
for wl in pypic explicit boris; do
CMD="python bench.py --workload $wl --steps 16 --warmup 3"
$CMD > gpurun_out/plain_$wl.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2_$wl.csv $CMD > gpurun_out/ncu_l_$wl.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2_$wl.csv | head -24
done
