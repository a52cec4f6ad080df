#!/usr/bin/env python
"""Peer-memory reduction (SheathSim(reduce="p2p")) against the NCCL all-reduce path on the same
shards: iteration counts and absorb flags identical, fields and particles equal to round-off.  (Not
bit-identical even at 2 ranks: the default build merges its deposit windows with fp64 REDs whose
order follows scheduling, so two NCCL runs differ from each other by the same few 1e-16 -- the
check runs the NCCL path twice and reports that noise floor next to the p2p difference.)
    torchrun --nproc-per-node 2 tools/p2p_check.py [N] [steps]"""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200.dist import Comm
from pypic_b200.sheath import SheathSim

KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27


def main():
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    Ng = 257; dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1)
    kT = KB * 116000.
    rs = np.random.RandomState(3)
    h = N // 2
    x0 = rs.uniform(0, L, N)
    u0 = np.concatenate([rs.normal(0, np.sqrt(kT / ME), h), rs.normal(0, np.sqrt(kT / MP), N - h)])
    E0 = rs.normal(0, 1e4, Ng)

    def run(reduce):
        sim = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=False, comm=Comm(), device=dev, rng="philox",
                        seed=7, sort_every=0, reduce=reduce)     # no sort: the counting sort's order inside a cell is not
                                                                 # reproducible and the Philox draws are keyed by slot
        sim.upload(x0, u0, E0=E0)
        its = []
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for s in range(steps):
            if s == steps - 2:
                sim._prev_hist = [1e30] * (len(sim._prev_hist) + 2)      # no-op launches + the repair pass (pic_dev_p2p_reduce)
            its.append(sim.step()[0])
        torch.cuda.synchronize(); t1 = time.perf_counter()
        sim.check()
        return sim, its, t1 - t0
    def rel(p, q):
        p, q = p.cpu().numpy(), q.cpu().numpy()
        return float(np.max(np.abs(p - q)) / max(np.max(np.abs(p)), 1e-300))
    a, its_a, ta = run("nccl")
    a2, its_a2, _ = run("nccl")
    out = dict(world=world, N=N, steps=steps)
    try:
        b, its_b, tb = run("p2p")
        out.update(iters_nccl=its_a, iters_p2p=its_b, repairs=(a.u_repairs, b.u_repairs),
                   E_rel=rel(a.E0, b.E0), j_rel=rel(a.j0, b.j0), x_rel=rel(a.x0, b.x0), u_rel=rel(a.u0, b.u0),
                   flags_equal=bool(torch.equal(a.active, b.active)),
                   nccl_vs_nccl=dict(E_rel=rel(a.E0, a2.E0), j_rel=rel(a.j0, a2.j0), x_rel=rel(a.x0, a2.x0)),
                   seconds=(ta, tb), seq=b.p2p.seq)
        out["ok"] = bool(its_a == its_b == its_a2 and out["flags_equal"] and out["E_rel"] < 1e-13 and out["j_rel"] < 1e-13
                         and out["x_rel"] < 1e-13 and out["u_rel"] < 1e-12)
        b.close()
    except Exception as e:                       # report instead of hanging the other rank's collectives
        out.update(ok=False, error=repr(e))
    allout = [None] * world
    dist.all_gather_object(allout, out)
    if rank == 0:
        print(json.dumps(dict(ranks=allout, ok=all(o["ok"] for o in allout))))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
