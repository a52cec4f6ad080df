# Round 2: the cell-sorted fast path behind the reference's own API (PIC_L_DD.main_i: host MT19937 draws in
# original-index order, carried v,w) -- parity tests, then the unchanged launcher shape at 2e7 particles against
# bench mode at the same size.
python -m pytest tests/test_gpu_sheath.py tests/test_gpu_dropin.py tests/test_checkpoint.py -m gpu -x -q 2>&1 | tail -5
mkdir -p /tmp/drv/plots && cd /tmp/drv
printf "import PIC_L_DD as p\nimport convert as c\n\ndef main():\n\tstart = 0\n\tstop = 1000\n\tskip = 10\n\tp.main_i(stop,skip)\n\nif __name__ == '__main__':\n\tmain()\n" > run_pypic_dd.py
for se in 8 0; do
PIC_TIMING=1 python $GRAFT_REPO_ROOT/tools/drive.py run_pypic_dd.py --seed 1 --set N=20000000 --set Ng=4097 --set sort_every=$se --steps 200 --quiet 2>&1 | grep -v Warning | tail -3
done
cd $GRAFT_REPO_ROOT
python bench.py --particles-per-gpu 2e7 --steps 200 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('bench mode 2e7:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'k', d['config']['picard_iterations_per_step'])"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_try1.json 2> gpurun_out/bench_r2_try1.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_try1.json')); print('%.3e'%d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step']); print(d['e2e']); print(d['strong_scaling']); print(d['cpu_baseline'])"; tail -3 gpurun_out/bench_r2_try1.err
