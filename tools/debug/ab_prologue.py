"""A/B on one box: SheathSim.step() with the merged prologue launch on and off, bench mode (device Philox)
and the reference-API mode (host MT19937 draws, tracked order, fused moments).  usage: ab_prologue.py [N] [steps]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pypic_b200.sheath import SheathSim
KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
Ng = 4097; dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1); kT = KB * 116000.


def make(api):
    np.random.seed(1)
    sim = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=api, rng="host" if api else "philox", seed=1,
                    sort_every=8, vion_after=2000 if api else None)
    g = torch.Generator(device=sim.dev); g.manual_seed(5)
    sim.x0.uniform_(0, 1, generator=g).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
    sim.u0.normal_(0, 1, generator=g); sim.u0[:N // 2].mul_(float(np.sqrt(kT / ME))); sim.u0[N // 2:].mul_(float(np.sqrt(kT / MP)))
    sim.fused_moments = api
    return sim


for api in (False, True):
    for rep in range(2):
        for merge in (False, True):
            sim = make(api)
            sim.merge_prologue = merge
            with sim.draws.hold():
                for _ in range(9):
                    sim.step()
                torch.cuda.synchronize(); t0 = time.perf_counter()
                for _ in range(steps):
                    sim.step()
                    if api:
                        sim.pre_step_moments(); sim.step_stats()
                torch.cuda.synchronize(); t = time.perf_counter() - t0
            print("%-14s merge=%-5s %.3f ms/step  (k=%d)" % ("reference API" if api else "bench mode", merge, 1e3 * t / steps, sim.last_iters), flush=True)
            sim.close() if hasattr(sim, "close") else None
            del sim; torch.cuda.empty_cache()
