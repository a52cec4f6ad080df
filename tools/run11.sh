set -x
timeout 600 python tools/kbench.py 2e8 window,window-ldg > gpurun_out/kbench1.json 2> gpurun_out/kbench1.err; tail -5 gpurun_out/kbench1.err
