#!/usr/bin/env python
"""Residual of every Picard iteration of the bench workload (how regular is the contraction that
SheathSim._expect_last relies on?) and which iterations ran as full ones."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pypic_b200.sheath import SheathSim

class A: pass
a = A(); a.particles_per_gpu = float(sys.argv[1]) if len(sys.argv) > 1 else 2e8; a.total_particles = 0; a.cells = 4096
w = bench.workload(a, 1)
dev = torch.device("cuda", 0)
sim = SheathSim(w["N"], w["Ng"], w["dx"], w["dt"], w["p2c"], tol=w["tol"], maxiter=w["maxiter"],
                kBT=(w["kBTe"], w["kBTi"]), carry_vw=False, rng="philox", seed=1, device=dev, sort_every=8)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
sim.x0.uniform_(0.0, 1.0, generator=gen).mul_(w["L"]).clamp_(1e-12, w["L"] * (1 - 1e-12))
sim.u0.normal_(0.0, 1.0, generator=gen)
sim.u0[:sim.n_split].mul_(float(np.sqrt(w["kBTe"] / bench.ME))); sim.u0[sim.n_split:].mul_(float(np.sqrt(w["kBTi"] / bench.MP)))
out = []
for step in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    sim.resid_trace = []; sim.iter_events = []
    sim.step()
    torch.cuda.synchronize()
    out.append({"r": sim.resid_trace, "full": [bool(e[2]) for e in sim.iter_events],
                "ms": [round(e[0].elapsed_time(e[1]), 4) for e in sim.iter_events]})
print(json.dumps({"tol": w["tol"], "steps": out, "repairs": sim.u_repairs}))
