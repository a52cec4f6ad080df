# wide-window build of the sheath kernel selected at run time: parity suites, then A/B against the 7-node build
python -m pytest tests/test_gpu_sheath.py tests/test_gpu_reproducible.py tests/test_gpu_dropin.py tests/test_checkpoint.py -x -q -m gpu 2>&1 | tail -4
run() {  # label env sort_every
  env $2 python bench.py --steps 48 --warmup 3 --sort-every $3 --no-e2e --no-cpu-baseline --strong-total 0 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); r=d['roofline']; a=d['reference_api']; print('$1 sort every $3:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'kernel ms %.4f'%r['kernel_ms_mean'], 'share %.3f'%r['kernel_share_of_step'], 'frac %.3f'%r['frac'], 'api %.3f ms'%a['ms_per_step'], d['clocks']['sm_mhz'])"
}
for rep in 1 2; do
run "narrow (7 nodes, 4 stages)" PIC_V6_NARROW=1 8
run "wide (15 nodes, 3 stages)" PIC_V6_NARROW=0 8
run "wide (15 nodes, 3 stages)" PIC_V6_NARROW=0 12
done
