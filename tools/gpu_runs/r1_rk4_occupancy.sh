for m in 4 5 6; do PIC_RK4_MINB=$m python tools/bench_paths.py 1e8 gc 4 1 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); r=[x for x in d['results'] if 'rk4' in x['path']][0]; print('minb $m', r['ms'])"; done
