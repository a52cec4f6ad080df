#!/usr/bin/env python
"""Host wall-clock and device-event time of the sections of SlabSheathSim.step() under a few variants (diagnostics)."""
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
for _ in range(3):
    sim.step()
ev = lambda: torch.cuda.Event(enable_timing=True)
for variant in ("ahead", "sync_before_step", "no_ahead"):
    sim.enqueue_ahead = variant != "no_ahead"
    torch.cuda.synchronize()
    host = {"reinject": 0.0, "migrate_sort": 0.0, "picard": 0.0}
    sec = {"reinject": [], "migrate_sort": [], "picard": []}
    w0 = time.perf_counter()
    for _ in range(steps):
        if variant == "sync_before_step":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        a = ev(); a.record(); sim.reinject(); sim._reset_logs(); b = ev(); b.record()
        t1 = time.perf_counter()
        if sim.sort_every and sim.t % sim.sort_every == 0:
            sim.migrate_sort()
        t2 = time.perf_counter()
        c = ev(); c.record(); sim.picard(); d = ev(); d.record(); sim.t += 1
        t3 = time.perf_counter()
        host["reinject"] += t1 - t0; host["migrate_sort"] += t2 - t1; host["picard"] += t3 - t2
        sec["reinject"].append((a, b)); sec["migrate_sort"].append((b, c)); sec["picard"].append((c, d))
    torch.cuda.synchronize()
    wall = (time.perf_counter() - w0) / steps * 1e3
    out = {"variant": variant, "wall_ms_per_step": round(wall, 3),
           "device_ms": {k: round(float(np.sum([x.elapsed_time(y) for x, y in v])) / steps, 3) for k, v in sec.items()},
           "host_ms": {k: round(1e3 * v / steps, 3) for k, v in host.items()}, "rank": rank}
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
