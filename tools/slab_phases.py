#!/usr/bin/env python
"""Device-time breakdown of SlabSheathSim.step() with the distributed field update (torchrun; diagnostics):
CUDA events around re-injection / sort+migration / Picard loop, and at the phase boundaries of every iteration
(particle kernels + pack | all-gather 1 | field kernel | all-gather 2 | finish).  No extra synchronisation."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200.dist import Comm
from pypic_b200.spatial import SlabSheathSim
KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(float(sys.argv[1])) * world; Ng = int(sys.argv[2]) if len(sys.argv) > 2 else 4097
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dx, dt = 1e-5, 1e-12; L = dx * (Ng - 1); kT = KB * 116000.
sim = SlabSheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), comm=Comm(), device=dev, sort_every=8,
                    field=os.environ.get("PIC_SLAB_FIELD", "distributed"))
sim.init_device(1234)
for _ in range(3):
    sim.step()
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
sec = {"reinject": [], "migrate_sort": [], "picard": []}
sim.phase_events = []
e_all0 = ev(); e_all0.record()
for _ in range(steps):
    a = ev(); a.record(); sim.reinject(); sim._reset_logs(); b = ev(); b.record()
    if sim.sort_every and sim.t % sim.sort_every == 0:
        sim.migrate_sort()
    c = ev(); c.record(); sim.picard(); d = ev(); d.record(); sim.t += 1
    sec["reinject"].append((a, b)); sec["migrate_sort"].append((b, c)); sec["picard"].append((c, d))
e_all1 = ev(); e_all1.record()
torch.cuda.synchronize()
out = {k: round(float(np.sum([x.elapsed_time(y) for x, y in v])) / steps, 3) for k, v in sec.items()}
out["step_ms"] = round(e_all0.elapsed_time(e_all1) / steps, 3)
pe = sim.phase_events
names = ["particles+pack", "all_gather_1", "field", "all_gather_2", "finish"]
if sim.field == "distributed" and len(pe) % 6 == 0:
    ph = {n: [] for n in names}
    for i in range(0, len(pe), 6):
        for j, n in enumerate(names):
            ph[n].append(pe[i + j].elapsed_time(pe[i + j + 1]))
    out["per_iteration_ms"] = {n: round(float(np.mean(v)), 4) for n, v in ph.items()}
    out["iterations_launched_per_step"] = len(pe) / 6 / steps
# second pass: synchronising wall-clock sections of the re-injection
sim.phase_events = None; sim.rprof = {}
for _ in range(steps):
    sim.step()
out["reinject_sections_ms"] = {k: round(1e3 * v / steps, 3) for k, v in sim.rprof.items()}
out["local_dead_last"] = sim.local_dead
out["rank"] = rank; out["world"] = world; out["N_per_rank"] = sim.local_particles(); out["Ng"] = Ng
print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
