# sheath window kernel: deposit-window width (V6_W, libraries prebuilt with -DV6_W=..) vs re-sort interval
for w in 7 9 11 7 9 11; do
  for se in 8 16; do
    LIBP=pypic_b200/_variants/libpic_b200_w$w.so; [ $w = 7 ] && LIBP=pypic_b200/libpic_b200.so
    PIC_LIB_PATH=$PWD/$LIBP python bench.py --steps 48 --warmup 3 --sort-every $se --no-e2e --no-cpu-baseline --no-api-leg --strong-total 0 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('V6_W=$w sort every $se:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'kernel ms %.4f'%r['kernel_ms_mean'], 'share %.3f'%r['kernel_share_of_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
