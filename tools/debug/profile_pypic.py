import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pypic_b200.periodic import PeriodicImplicitSim
from pypic_b200 import _lib, device as D
KB, ME = 1.38E-23, 9.11E-31
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000000
cells = 4096; dx, dt = 1e-5, 1e-12; L = dx * cells; kT = KB * 116000.
dev = torch.device("cuda", 0)
sim = PeriodicImplicitSim(N, cells, dx, dt, L, L * 1e19 / N, tol=1e-3, maxiter=20, device=dev, sort_every=8, track_order=False)
g = torch.Generator(device=dev); g.manual_seed(1)
sim.x0.uniform_(0., 1., generator=g).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
sim.v0.normal_(0., 1., generator=g).mul_(float(np.sqrt(kT / ME)))
for _ in range(4): sim.push()
torch.cuda.synchronize()
# count syncs: wrap D.read_f64
calls = {"read": 0, "call": 0}
orig_read = D.read_f64
def rd(*a, **k):
    calls["read"] += 1; return orig_read(*a, **k)
D.read_f64 = rd
import pypic_b200.periodic as P
P.D.read_f64 = rd
t0 = time.perf_counter(); ks = []
for s in range(24):
    ks.append(sim.push()[0])
torch.cuda.synchronize(); t1 = time.perf_counter()
print("steps 24: %.3f ms/step, iterations %s, host reads per step %.2f, repairs %d, prev_hist %s" % (1e3 * (t1 - t0) / 24, ks[:8], calls["read"] / 24, sim.j1_repairs, sim._prev_hist))
# per-step GPU time by events without per-iteration events
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(24): sim.push()
e1.record(); torch.cuda.synchronize()
print("event-timed: %.3f ms/step" % (e0.elapsed_time(e1) / 24))
os.environ["X"] = "1"
