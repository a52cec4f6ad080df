import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200.sheath import SheathSim
Ng = 4097; dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1)
KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
kT = KB * 116000.
dev = torch.device("cuda", 0)
N = int(float(sys.argv[1])); dep = sys.argv[2]; se = int(sys.argv[3]); nst = int(sys.argv[4]) if len(sys.argv) > 4 else 12
sim = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=False, deposit=dep, rng="philox", seed=1, device=dev, sort_every=se)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
sim.x0.uniform_(0.0, 1.0, generator=gen).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
sim.u0.normal_(0.0, 1.0, generator=gen)
sim.u0[:sim.n_split].mul_(float(np.sqrt(kT / ME))); sim.u0[sim.n_split:].mul_(float(np.sqrt(kT / MP)))
ks = []
for st in range(nst):
    sim.resid_trace = []
    sim.step()
    tr = sim.resid_trace
    bad = any(tr[i + 1] > 0.05 * tr[i] for i in range(len(tr) - 1))
    ks.append(len(tr))
    if bad:
        print("N %.0e %s se=%d step %d KICK" % (N, dep, se, st), " ".join("%.1e" % r for r in tr))
print("N %.0e %s se=%d iterations per step:" % (N, dep, se), ks)
