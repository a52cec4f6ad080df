"""Where a step of the reference-API path (host MT19937 draws, tracked order) spends its time:
wall-clock per section with a device synchronize after each (so GPU time is attributed to its section)."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pypic_b200.sheath import SheathSim
from pypic_b200.rng import LegacyDraws
KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sync = (sys.argv[3] != "nosync") if len(sys.argv) > 3 else True
Ng = 4097; dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1); kT = KB * 116000.
np.random.seed(1)
h = N // 2
sim = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=True, rng="host", sort_every=8, vion_after=2000)
g = torch.Generator(device=sim.dev); g.manual_seed(5)
sim.x0.uniform_(0, 1, generator=g).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
sim.u0.normal_(0, 1, generator=g); sim.u0[:h].mul_(float(np.sqrt(kT / ME))); sim.u0[h:].mul_(float(np.sqrt(kT / MP)))
T = {}
def sec(name, fn):
    t0 = time.perf_counter(); out = fn()
    if sync: torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return out
for _ in range(5):
    sim.step()
T.clear()
torch.cuda.synchronize(); t_all = time.perf_counter()
sim.fused_moments = True
with sim.draws.hold():
    for s in range(steps):
        sec("reinject", sim.reinject)
        if sim.t % sim.sort_every == 0:
            sec("sort", sim.sort_by_cell)
        sec("picard", sim.picard); sim.t += 1
        sim.pre_step_moments(); sim.step_stats()
torch.cuda.synchronize(); t_all = time.perf_counter() - t_all
print("N=%g sync=%s: %.3f ms/step;" % (N, sync, 1e3 * t_all / steps), {k: "%.3f" % (1e3 * v / steps) for k, v in T.items()},
      "jumps", sim.draws.jumps, "prefetched", sim.draws.prefetch_hits, "k", sim.last_iters)
# finer: inside reinject
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
with sim.draws.hold():
    for s in range(50):
        sim.step(); sim.pre_step_moments()
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(28); print(st.getvalue()[:5000])
