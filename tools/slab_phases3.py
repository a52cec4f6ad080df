#!/usr/bin/env python
"""Per-step host time of SlabSheathSim.reinject() / picard() over the first steps of a run (diagnostics)."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200.dist import Comm
from pypic_b200.spatial import SlabSheathSim
KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(float(sys.argv[1])) * world; Ng = int(sys.argv[2]); steps = int(sys.argv[3])
dx, dt = 1e-5, 1e-12; L = dx * (Ng - 1); kT = KB * 116000.
sim = SlabSheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), comm=Comm(), device=dev, sort_every=8)
sim.init_device(1234)
rows = []
for s in range(steps):
    sim.rprof = {} if s >= steps - 6 and s < steps - 3 else None      # three late steps with synchronising sections
    t0 = time.perf_counter(); sim.reinject(); sim._reset_logs(); t1 = time.perf_counter()
    if sim.sort_every and sim.t % sim.sort_every == 0:
        sim.migrate_sort()
    t2 = time.perf_counter(); k, r = sim.picard(); sim.t += 1; t3 = time.perf_counter()
    rows.append([round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t1), 2), round(1e3 * (t3 - t2), 2), list(sim.local_dead), sim.stat["exported"]])
if rank == 0:
    for s, r_ in enumerate(rows):
        print(s, r_, flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
