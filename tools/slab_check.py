#!/usr/bin/env python
"""Slab (spatial) decomposition parity check, run under torchrun (one rank per GPU):
SlabSheathSim on W ranks against the SAME global particles on one rank (no process group), and
its first step against the particle-decomposed SheathSim.  Re-injection draws are keyed by the
global ordinal of the dead particle, so the particle SET is identical for any W: sorted
positions / velocities and the fields must agree to round-off.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200.dist import Comm  # noqa: E402
from pypic_b200.sheath import SheathSim  # noqa: E402
from pypic_b200.spatial import SlabSheathSim  # noqa: E402

KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400000
    Ng = int(sys.argv[2]) if len(sys.argv) > 2 else 513
    steps = 7
    dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1)
    kT = KB * 116000.
    rs = np.random.RandomState(3)
    h = N // 2
    x0 = rs.uniform(0, L, N)
    u0 = np.concatenate([rs.normal(0, np.sqrt(kT / ME), h), rs.normal(0, np.sqrt(kT / MP), N - h)])
    E0 = rs.normal(0, 1e4, Ng)
    p2c = L * 1e19 / N

    field = os.environ.get("PIC_SLAB_FIELD", "distributed")

    def run(comm):
        sim = SlabSheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), comm=comm, device=dev, sort_every=2, guard=16, seed=5,
                            field=field)
        sim.upload(x0, u0, E0)
        its, counts = [], []
        for _ in range(steps):
            k, r = sim.step()
            its.append(k)
            counts.append(sim.local_particles())
        sim.check()
        return sim, its, counts
    slab, its_s, cnt_s = run(Comm())
    parts = slab.gather_particles()
    E_slab = slab.gather_field("E0").cpu().numpy()        # every rank holds its own nodes + guard nodes only
    tot = torch.tensor([slab.local_particles()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    # ownership invariant right after a migration: every particle inside the slab +- guard
    inside = True
    for blk in slab.blocks:
        x = blk.x0[:blk.n]
        cell = torch.floor(x / dx)
        act = blk.active[:blk.n] == 1
        inside = inside and bool(((cell[act] >= slab.c0 - slab.G) & (cell[act] < slab.c1 + slab.G)).all().item())
    if rank == 0:
        one, its_1, cnt_1 = run(Comm(enabled=False))
        ref = one.gather_particles()
        res = dict(world=world, N=N, Ng=Ng, field=field, iters_slab=its_s, iters_single=its_1, total_particles=int(tot.item()),
                   local_counts_rank0=cnt_s, stat=slab.stat, inside_guard=inside,
                   E_rel=rel(E_slab, one.E0.cpu().numpy()))
        ok = its_s == its_1 and int(tot.item()) == N and inside and res["E_rel"] < 1e-9
        for name, i in (("electrons", 0), ("ions", 3)):
            xa, xb = np.sort(parts[i]), np.sort(ref[i])
            res[name + "_x_rel"] = rel(xa, xb) if len(xa) == len(xb) else None
            res[name + "_dead"] = [int((parts[i + 2] != 1).sum()), int((ref[i + 2] != 1).sum())]
            ok = ok and len(xa) == len(xb) and res[name + "_x_rel"] < 1e-10 and res[name + "_dead"][0] == res[name + "_dead"][1]
        # first step against the particle-decomposed path (no re-injection has happened yet)
        a = SlabSheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), comm=Comm(enabled=False), device=dev, sort_every=0, seed=5)
        a.upload(x0, u0, E0); ka, _ = a.picard()
        b = SheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), carry_vw=False, comm=Comm(enabled=False), device=dev)
        b.upload(x0, u0, E0=E0); kb_, _ = b.picard()
        res["first_step_E_rel_vs_particle_decomposition"] = rel(a.E0.cpu().numpy(), b.E0.cpu().numpy())
        ok = ok and ka == kb_ and res["first_step_E_rel_vs_particle_decomposition"] < 1e-12
        res["ok"] = bool(ok)
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
