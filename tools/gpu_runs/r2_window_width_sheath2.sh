# sheath window kernel: wider windows at 3 ring stages, and the re-sort interval in between
run() {  # lib label sort_every
  PIC_LIB_PATH=$PWD/$1 python bench.py --steps 48 --warmup 3 --sort-every $3 --no-e2e --no-cpu-baseline --no-api-leg --strong-total 0 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('$2 sort every $3:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'kernel ms %.4f'%r['kernel_ms_mean'], 'share %.3f'%r['kernel_share_of_step'], d['clocks']['sm_mhz'])"
}
for rep in 1 2; do
run pypic_b200/libpic_b200.so "V6_W=7  NST=4" 8
run pypic_b200/_variants/libpic_b200_w11.so "V6_W=11 NST=4" 8
run pypic_b200/_variants/libpic_b200_w11.so "V6_W=11 NST=4" 12
run pypic_b200/_variants/libpic_b200_w11n3.so "V6_W=11 NST=3" 8
run pypic_b200/_variants/libpic_b200_w13n3.so "V6_W=13 NST=3" 8
run pypic_b200/_variants/libpic_b200_w13n3.so "V6_W=13 NST=3" 12
run pypic_b200/_variants/libpic_b200_w15n3.so "V6_W=15 NST=3" 8
run pypic_b200/_variants/libpic_b200_w15n3.so "V6_W=15 NST=3" 12
run pypic_b200/_variants/libpic_b200_w15n3.so "V6_W=15 NST=3" 16
done
