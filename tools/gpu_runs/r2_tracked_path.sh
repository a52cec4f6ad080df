# Round 2: the cell-sorted fast path behind the reference's own API (PIC_L_DD.main_i: host MT19937 draws in
# original-index order, carried v,w) -- the whole GPU suite, then the unchanged launcher shape at 2e7 particles
# against bench mode at the same size.
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
mkdir -p /tmp/drv/plots && cd /tmp/drv
printf "import PIC_L_DD as p\nimport convert as c\n\ndef main():\n\tstart = 0\n\tstop = 1000\n\tskip = 10\n\tp.main_i(stop,skip)\n\nif __name__ == '__main__':\n\tmain()\n" > run_pypic_dd.py
for n in 20000000 200000000; do
PIC_TIMING=1 python $GRAFT_REPO_ROOT/tools/drive.py run_pypic_dd.py --seed 1 --set N=$n --set Ng=4097 --steps 200 --quiet 2>&1 | grep -v Warning | tail -3
done
cd $GRAFT_REPO_ROOT
for n in 2e7 2e8; do
python bench.py --particles-per-gpu $n --steps 200 --warmup 3 --no-e2e --no-cpu-baseline --strong-total 0 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('bench mode $n:', '%.3e'%d['value'], '%.3f ms/step'%d['ms_per_step'], 'k', d['config']['picard_iterations_per_step'])"
done
