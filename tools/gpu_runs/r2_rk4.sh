# GC RK4 (pygcpic.py:598-645): the pair kernel against the one-particle-per-thread kernel, then an ncu capture
python -m pytest tests/test_gpu_gc.py -m gpu -x -q 2>&1 | tail -3
for pr in 0 2 3 4; do
PIC_RK4_PAIR=$pr python tools/bench_paths.py 1e8 gc 4 1 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for r in d['results']:
    if 'rk4' in r.get('path',''): print('PIC_RK4_PAIR=$pr', r['path'], '%.3f ms'%r['ms'], '%.3e p-s/s'%r['particle_steps_per_s'], 'frac(64B) %.3f'%r['frac'])"
done
CMD="python tools/bench_paths.py 1e8 gc 2 1"
$CMD > gpurun_out/plain_rk4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gc_push_rk4_uniform -s 1 -c 1 -o gpurun_out/prof_r2_rk4 $CMD > gpurun_out/ncu_rk4.log 2>&1
tail -2 gpurun_out/ncu_rk4.log
