set -x
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --particles-per-gpu 1e8"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dd_picard_iter_win -s 6 -c 3 -o gpurun_out/prof_r1_iter $CMD > gpurun_out/ncu.log 2>&1
tail -5 gpurun_out/ncu.log
ls -la gpurun_out
