# A/B of the enqueue-ahead Picard loops on one box (PIC_ENQUEUE_AHEAD=0: one host round trip per iteration)
for wl in pypic sheath; do for a in 0 1 0 1; do
PIC_ENQUEUE_AHEAD=$a python bench.py --workload $wl --steps 40 --warmup 4 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$wl ahead=$a', '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'kernel %.4f'%d['roofline']['kernel_ms_mean'], 'share %.3f'%d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done; done
