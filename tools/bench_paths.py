#!/usr/bin/env python
"""Per-path micro-benchmarks of the non-sheath movers of SURVEY.md 8(d): explicit leapfrog
(PIC_L, 32 B/particle-step), periodic CN/Picard (pypic, 32k+16), Boris 1D3V + n,rho deposit
(pygcpic, 64 B carried minimum) and GC RK4.  CUDA events, inputs larger than L2, one JSON
object on stdout.  Usage: bench_paths.py [N] [paths,comma-separated] [reps]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200 import _lib, device as D  # noqa: E402
from pypic_b200.periodic import ExplicitSim, PeriodicImplicitSim  # noqa: E402
from pypic_b200.gcstore import GridDev, ParticleStore  # noqa: E402

KB, ME, MP, E_CH = 1.38E-23, 9.11E-31, 1.67E-27, 1.602E-19
PEAK = 6555.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps, warm=2, prep=None):
    for _ in range(warm):
        if prep:
            prep()
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if prep:
            prep()                       # untimed: restore the inputs (e.g. the freshly sorted positions)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts


def line(name, N, ts, alg_bytes, extra=None):
    ms = float(np.mean(ts))
    d = dict(path=name, N=N, ms=ms, ms_min=float(np.min(ts)), particle_steps_per_s=N / (ms * 1e-3),
             alg_bytes_per_particle=alg_bytes, achieved_gbs=N * alg_bytes / (ms * 1e-3) / 1e9,
             frac=N * alg_bytes / (ms * 1e-3) / 1e9 / PEAK)
    if extra:
        d.update(extra)
    return d


def bench_explicit(N, reps, dev, sort):
    Ng = 4096; dx = 1e-5; dt = 1e-12
    L = dx * (Ng - 1)
    kT = KB * 116000.
    sim = ExplicitSim(N, Ng, dx, dt, (L + dx) * 1e19 / N, q=(-E_CH, E_CH), m=(ME, MP), n_split=N // 2, device=dev,
                      sort_every=8 if sort else 0)
    g = torch.Generator(device=dev); g.manual_seed(7)
    sim.x.uniform_(0., 1., generator=g).mul_(L + dx).clamp_(1e-12, (L + dx) * (1 - 1e-12))
    if sort:
        h = N // 2
        sim.x[:h] = torch.sort(sim.x[:h])[0]; sim.x[h:] = torch.sort(sim.x[h:])[0]
    sim.v.normal_(0., 1., generator=g)
    sim.v[:N // 2].mul_(float(np.sqrt(kT / ME))); sim.v[N // 2:].mul_(float(np.sqrt(kT / MP)))
    sim.step()
    ts_push = timed(sim.push, reps)
    ts_step = timed(sim.step, reps)
    sim.check()
    return [line("explicit_push_deposit(sorted=%d)" % sort, N, ts_push, 32.0),
            line("explicit_full_step(sorted=%d)" % sort, N, ts_step, 32.0)]


def bench_pypic(N, reps, dev, sort):
    Ng = 4096; dx = 1e-5; dt = 1e-12
    L = dx * Ng
    kT = KB * 116000.
    sim = PeriodicImplicitSim(N, Ng, dx, dt, L, L * 1e19 / N, device=dev, tol=1e-3, sort_every=8 if sort else 0)
    g = torch.Generator(device=dev); g.manual_seed(7)
    sim.x0.uniform_(0., 1., generator=g).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
    if sort:
        sim.x0.copy_(torch.sort(sim.x0)[0])
    sim.v0.normal_(0., 1., generator=g).mul_(float(np.sqrt(kT / ME)))
    P = C.byref(sim.params)
    st = D.stream()
    sim.Es.zero_(); sim.Fs.zero_()

    def it(first):
        def f():
            _lib.call("pic_dev_pypic_picard_iter", P, D.ptr(sim.x0), D.ptr(sim.v0), D.ptr(sim.x1), D.ptr(sim.v1),
                      D.ptr(sim.Fs), D.ptr(sim.acc), first, D.ptr(sim.range_err), st)
        return f
    t1 = timed(it(1), reps)
    t0 = timed(it(0), reps)
    ks = []

    def push():
        k, r = sim.push(); ks.append(k)
    tp = timed(push, reps)
    sim.check()
    k = float(np.mean(ks))
    return [line("pypic_picard_iter_first(sorted=%d)" % sort, N, t1, 32.0),
            line("pypic_picard_iter_later(sorted=%d)" % sort, N, t0, 40.0),
            line("pypic_push_full(sorted=%d)" % sort, N, tp, 32. * k + 16., dict(picard_iterations=k))]


def bench_gc(N, reps, dev, sort):
    ng = 4097
    Te = 60. * 11600.; Ti = 50. * 11600.
    lamD = np.sqrt(8.854e-12 * KB * Te / (1e19 * E_CH ** 2))
    Lg = 100. * lamD * (ng - 1) / 149.      # the reference's resolution (ng=150 over 100 Debye lengths) at 4097 nodes
    alpha = 86. * np.pi / 180.
    grid = GridDev(ng, Lg, Te, device=dev)
    store = ParticleStore(N, B=(2. * np.cos(alpha), 2. * np.sin(alpha), 0.), device=dev)
    g = torch.Generator(device=dev); g.manual_seed(7)
    store.r[0].uniform_(0., 1., generator=g).mul_(Lg).clamp_(Lg * 1e-9, Lg * (1 - 1e-9))
    if sort:
        store.r[0].copy_(torch.sort(store.r[0])[0])
    vth = float(np.sqrt(KB * Ti / MP))
    for c in (3, 4, 5):
        store.r[c].normal_(0., vth, generator=g)
    store.charge_state.fill_(1.); store.m.fill_(MP); store.p2c.fill_(Lg * 1e19 / N); store.Z.fill_(1)
    grid.E.normal_(0., 1e3, generator=g)
    dt = 1e-10
    x_keep = store.r[0].clone()
    out = []

    def restore():
        store.r[0].copy_(x_keep); store.active.fill_(1); store.at_wall.fill_(0)

    def boris():
        store.push_6D(dt, grid)
    ts = timed(boris, reps, prep=restore)
    out.append(line("gc_push_boris(gather+push+bc)(sorted=%d)" % sort, N, ts, 112.0))

    def boris_dep():
        store.push_6D(dt, grid, deposit=True)
        grid.finish_fused_deposit(1.0, dt)
    ts = timed(boris_dep, reps, prep=restore)
    out.append(line("gc_push_boris+deposit_fused(sorted=%d)" % sort, N, ts, 112.0))
    store.FUSED_MIN = 10 ** 12
    ts = timed(boris, reps, prep=restore)
    out.append(line("gc_push_boris_v1(sorted=%d)" % sort, N, ts, 112.0))
    store.FUSED_MIN = 16384
    restore()

    def weight():
        grid.weight_particles_to_grid_boltzmann(store, dt)
    ts = timed(weight, reps)
    out.append(line("gc_weight_n_rho(sorted=%d)" % sort, N, ts, 8.0))
    store.transform_6D_to_GC()
    store.r[3].abs_().clamp_(min=1.0)      # vpar != 0 (the reference's EOM is singular there)

    def rk4():
        store.push_GC(dt, grid)
    ts = timed(rk4, reps)
    out.append(line("gc_push_rk4(sorted=%d)" % sort, N, ts, 64.0))
    return out


def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
    paths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["explicit", "pypic", "gc"]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    sorts = [int(s) for s in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 0]
    dev = torch.device("cuda", 0)
    res = []
    for p in paths:
        for s in sorts:
            fn = dict(explicit=bench_explicit, pypic=bench_pypic, gc=bench_gc)[p]
            try:
                res += fn(N, reps, dev, s)
            except Exception as ex:       # keep the other paths' numbers
                res.append(dict(path=p, sorted=s, error=repr(ex)))
            torch.cuda.empty_cache()
    print(json.dumps(dict(peak_gbs=PEAK, results=res), indent=1))


if __name__ == "__main__":
    main()
