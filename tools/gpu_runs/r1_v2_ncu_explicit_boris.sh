set -x
for w in explicit boris; do
case $w in explicit) K=l_push_deposit_v2;; pypic) K=pypic_picard_iter_v2;; boris) K=gc_push_boris_v2;; esac
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -o gpurun_out/prof_r1_$w python bench.py --workload $w --steps 4 --warmup 3 > gpurun_out/ncu_$w.log 2>&1
tail -2 gpurun_out/ncu_$w.log
done
