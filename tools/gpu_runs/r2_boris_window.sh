# Boris: deposit-window width vs re-sort interval (ions cross ~0.6 cells per step at the reference's resolution)
for gw in 7 11 15; do
  PIC_NVCC_EXTRA="-DG_W=$gw" python -m pypic_b200.build > /dev/null 2>&1 || echo build failed
  for se in 8 16 32; do
    PIC_NVCC_EXTRA="-DG_W=$gw" python bench.py --workload boris --steps 48 --warmup 3 --sort-every $se 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('G_W=$gw sort every %d steps:'%($se//2), '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'kernel ms %.3f'%d['roofline']['kernel_ms_mean'], 'frac %.3f'%d['roofline']['frac'])"
  done
done
