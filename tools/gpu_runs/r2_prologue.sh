#!/bin/bash
# merged step prologue + tail folded into the v6 kernel + pinned reads: parity suites, gap profile, bench
python -m pytest tests/test_gpu_sheath.py tests/test_gpu_dropin.py tests/test_gpu_reproducible.py tests/test_gpu_init.py tests/test_checkpoint.py -x -q -m gpu 2>&1 | tail -8
python tools/debug/profile_step_gaps.py 2e8 17
python bench.py --steps 40 --warmup 4 --no-e2e --no-cpu-baseline --strong-total 0 > gpurun_out/bench_prologue.json 2> gpurun_out/bench_prologue.err; echo bench rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_prologue.json"))
print("value %.4e ms %.3f frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_share_of_step"]))
print("api", d.get("reference_api", {}).get("value"), d.get("reference_api", {}).get("ms_per_step"))
PY
