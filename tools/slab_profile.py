#!/usr/bin/env python
"""Wall-clock breakdown of SlabSheathSim.step() (torchrun, diagnostics)."""
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
dx, dt = 1e-5, 1e-12; L = dx * (Ng - 1); kT = KB * 116000.
sim = SlabSheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), comm=Comm(), device=dev, sort_every=8)
sim.init_device(1234)
for _ in range(3):
    sim.step()
sim.profile = {}
steps = 16
for _ in range(steps):
    sim.step()
# compaction alone
import time
from pypic_b200 import _lib, device as D
blk = sim.blocks[0]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    _lib.call("pic_dev_compact_flags", D.ptr(blk.active), blk.n, 0, D.ptr(blk.dead_idx), D.ptr(blk.count), D.ptr(blk.block_counts), D.stream())
torch.cuda.synchronize(); sim.profile["compact_flags_alone"] = (time.perf_counter() - t0) / 10 * steps
out = {k: round(1e3 * v / steps, 3) for k, v in sim.profile.items()}
out["rank"] = rank; out["stat"] = sim.stat
print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
