set -x
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --particles-per-gpu 1e8 --sort-every 8"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dd_picard_iter_v6 -s 16 -c 2 -o gpurun_out/prof_r1_v6 $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
