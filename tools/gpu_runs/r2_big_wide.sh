# large-grid build with the wide windows: parity (large-grid tests), then A/B at 1e6 cells (config 5's grid), one GPU
python -m pytest tests/test_gpu_sheath.py tests/test_gpu_reproducible.py tests/test_gpu_periodic.py -x -q -m gpu -k "big or large or grid" 2>&1 | tail -3
run() {
  env $2 python bench.py --steps 24 --warmup 3 --cells 1000000 --particles-per-gpu $3 --no-e2e --no-cpu-baseline --no-api-leg --strong-total 0 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('$1 N=$3:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'kernel ms %.4f'%r['kernel_ms_mean'], 'frac %.3f'%r['frac'], 'k %.1f'%d['config']['picard_iterations_per_step'], d['clocks']['sm_mhz'])"
}
for rep in 1 2; do
run "narrow" PIC_V6_NARROW=1 200000000
run "wide  " PIC_V6_NARROW=0 200000000
done
run "narrow" PIC_V6_NARROW=1 12500000
run "wide  " PIC_V6_NARROW=0 12500000
