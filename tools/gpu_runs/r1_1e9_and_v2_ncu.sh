set -x
timeout 600 python bench.py --total-particles 1e9 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_1e9_1gpu.json 2> gpurun_out/bench_r1_1e9_1gpu.err; tail -3 gpurun_out/bench_r1_1e9_1gpu.err; cat gpurun_out/bench_r1_1e9_1gpu.json
for w in explicit pypic boris; do
case $w in explicit) K=l_push_deposit_v2;; pypic) K=pypic_picard_iter_v2;; boris) K=gc_push_boris_v2;; esac
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -o gpurun_out/prof_r1_$w python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/ncu_$w.log 2>&1
tail -2 gpurun_out/ncu_$w.log
done
