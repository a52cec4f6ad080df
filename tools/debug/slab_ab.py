"""A/B of SlabSheathSim.step(): enqueue-ahead Picard loop on / off, wall clock per step (torchrun).
usage: slab_ab.py particles_per_rank [Ng] [steps]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pypic_b200.dist import Comm
from pypic_b200.spatial import SlabSheathSim
KB = 1.38E-23
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(float(sys.argv[1])) * world; Ng = int(sys.argv[2]) if len(sys.argv) > 2 else 4097
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dx, dt = 1e-5, 1e-12; L = dx * (Ng - 1); kT = KB * 116000.
comm = Comm()
for ahead in (False, True, False, True):
    sim = SlabSheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), comm=comm, device=dev, sort_every=8)
    sim.enqueue_ahead = ahead
    sim.init_device(1234)
    for _ in range(3):
        sim.step()
    torch.cuda.synchronize(); comm.barrier(); t0 = time.perf_counter()
    ks = [sim.step()[0] for _ in range(steps)]
    torch.cuda.synchronize(); comm.barrier(); t = time.perf_counter() - t0
    # sections (synchronising)
    sim.profile = {}
    for _ in range(8):
        sim.step()
    prof = {k: round(1e3 * v / 8, 3) for k, v in sim.profile.items()}
    if rank == 0:
        print("enqueue_ahead=%-5s %.3f ms/step  k=%.1f   sections (sync, ms/step): %s" % (ahead, 1e3 * t / steps, np.mean(ks), prof), flush=True)
    del sim; torch.cuda.empty_cache()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
