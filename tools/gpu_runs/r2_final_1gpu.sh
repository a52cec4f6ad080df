#!/bin/bash
# end of round 2: full GPU suite, smoke, default bench line, explicit workload with its new default sort interval
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_final.json')); r=d['roofline']; print('sheath %.3e %.3f ms/step frac %.3f share %.3f e2e %.3e api %.3e cpu %.3e launches %d' % (d['value'], d['ms_per_step'], r['frac'], r['kernel_share_of_step'], d['e2e']['value'], d['reference_api']['value'], d['cpu_baseline']['value'], d['gpu_launches']), d['clocks'])"
python bench.py --workload explicit --steps 48 --warmup 4 2>/dev/null > gpurun_out/bench_r2_final_explicit.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_final_explicit.json')); print('explicit %.3e %.3f ms/step frac %.3f sort_every %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['sort_every']))"
