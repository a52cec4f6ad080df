"""Where the sheath step's time goes besides the dominant kernel: CUDA events around every C-ABI call of a
few steps (bench configuration), printed as a timeline of one step -- call durations and the idle gaps
between them.  usage: profile_step_gaps.py [N] [steps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pypic_b200.sheath import SheathSim
from pypic_b200 import _lib
import pypic_b200.sheath as S
ME, MP, E = 9.11e-31, 1.67e-27, 1.602e-19
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
Ng = 4097; dx, dt = 1e-5, 1e-12; L = dx * (Ng - 1)          # bench.py's workload()
kBTe = kBTi = 1.38e-23 * 10.0 * 11600.
dev = torch.device("cuda", 0)
sim = SheathSim(N, Ng, dx, dt, L * 1e19 / N, tol=1e-5, maxiter=20, kBT=(kBTe, kBTi), carry_vw=False, deposit="window",
                rng="philox", seed=1, device=dev, sort_every=8)
gen = torch.Generator(device=dev); gen.manual_seed(1234)
sim.x0.uniform_(0., 1., generator=gen).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
sim.u0.normal_(0., 1., generator=gen)
sim.u0[:sim.n_split].mul_(float(np.sqrt(kBTe / ME))); sim.u0[sim.n_split:].mul_(float(np.sqrt(kBTi / MP)))
for _ in range(10):
    sim.step()
torch.cuda.synchronize()
log = []
orig = _lib.call


def call(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(name, *a); e1.record()
    log.append((name, e0, e1))
    return r


_lib.call = call
S._lib.call = call
marks = []
for s in range(steps):
    marks.append(len(log))
    sim.step()
marks.append(len(log))
torch.cuda.synchronize()
_lib.call = orig; S._lib.call = orig
# a step without a sort: print its timeline
tot = {}
gaps = []
for s in range(steps - 1):
    seg = log[marks[s]:marks[s + 1] + 1]           # + the first call of the next step (for the boundary gap)
    for i, (name, e0, e1) in enumerate(seg[:-1]):
        d = e0.elapsed_time(e1) * 1e3
        g = e1.elapsed_time(seg[i + 1][1]) * 1e3
        t = tot.setdefault(name, [0, 0.0, 0.0]); t[0] += 1; t[1] += d; t[2] += g
    gaps.append(seg[-2][2].elapsed_time(seg[-1][1]) * 1e3)
span = log[marks[0]][1].elapsed_time(log[marks[steps - 1]][1]) / (steps - 1)
print("N %d: %.3f ms/step over %d steps (events around every call: slightly slower than the bench)" % (N, span, steps - 1))
print("%-34s %6s %12s %12s" % ("call", "n/step", "us/step", "gap-after us/step"))
for name, (n, d, g) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %6.2f %12.1f %12.1f" % (name, n / (steps - 1), d / (steps - 1), g / (steps - 1)))
print("step-boundary gap (last call of a step -> first call of the next): mean %.1f us, max %.1f us" % (np.mean(gaps), np.max(gaps)))
s = steps // 2
print("timeline of step %d:" % s)
seg = log[marks[s]:marks[s + 1] + 1]
for i, (name, e0, e1) in enumerate(seg[:-1]):
    print("  %-34s %9.1f us   then idle %7.1f us" % (name, e0.elapsed_time(e1) * 1e3, e1.elapsed_time(seg[i + 1][1]) * 1e3))
