"""Per-step wall time and absorbed count of the reference-API path from a cold start (the first steps absorb
orders of magnitude more particles than the steady state).  usage: api_transient.py [N] [steps]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pypic_b200.sheath import SheathSim
from pypic_b200.rng import LegacyDraws
KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 45
Ng = 4097; dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1); kT = KB * 116000.
sim = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=True, rng="host", seed=1, sort_every=8, vion_after=2000,
                draws=LegacyDraws(np.random.RandomState(1)))
g = torch.Generator(device=sim.dev); g.manual_seed(4321)
sim.x0.uniform_(0, 1, generator=g).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
sim.u0.normal_(0, 1, generator=g); sim.u0[:N // 2].mul_(float(np.sqrt(kT / ME))); sim.u0[N // 2:].mul_(float(np.sqrt(kT / MP)))
sim.fused_moments = True
torch.cuda.synchronize()
rows = []
with sim.draws.hold():
    for s in range(steps):
        t0 = time.perf_counter()
        k, r = sim.step()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        nd = len(sim._harvested[1]) if sim._harvested is not None else -1
        rows.append((s, 1e3 * (t1 - t0), k, nd, sim.draws.jumps, sim.draws.prefetch_hits))
for r in rows:
    print("step %3d  %8.3f ms  k=%d  re-injected at its start %7d  jumps %d prefetched %d" % r)
