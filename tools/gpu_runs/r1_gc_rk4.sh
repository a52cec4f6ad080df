set -x
timeout 900 python -m pytest tests/test_gpu_gc.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python tools/bench_paths.py 1e8 gc 4 1 > gpurun_out/paths_gc3.json 2> gpurun_out/paths_gc3.err; tail -5 gpurun_out/paths_gc3.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/paths_gc3.json'))
for r in d['results']:
    if 'error' in r: print(r); continue
    print("%-55s %9.3f ms  %.3e p-s/s  %7.1f GB/s  frac %.3f" % (r['path'], r['ms'], r['particle_steps_per_s'], r['achieved_gbs'], r['frac']))
PY
