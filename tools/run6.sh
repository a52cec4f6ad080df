cp pypic_b200/libpic_b200.so /tmp/orig.so
for T in 512 640 768 896; do
  if [ $T != 512 ]; then cp pypic_b200/libpic_b200_T$T.so pypic_b200/libpic_b200.so; fi
  timeout 600 python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_T$T.json 2> gpurun_out/bench_T$T.err; tail -2 gpurun_out/bench_T$T.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_T$T.json')); print('T=$T', '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'])"
done
cp /tmp/orig.so pypic_b200/libpic_b200.so
