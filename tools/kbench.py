#!/usr/bin/env python
"""Micro-benchmark of the fused sheath Picard kernel on a sorted store: per-launch CUDA-event
times for the kernel variants and the per-CTA %globaltimer spread (load balance)."""
import ctypes as C
import json
import sys
import os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200 import _lib, device as D
from pypic_b200.sheath import SheathSim

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000000
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["window", "window-ldg"]
Ng = 4097; dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1)
KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
kT = KB * 116000.
dev = torch.device("cuda", 0)
out = {}
for var in variants:
    sim = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=False, deposit=var, rng="philox", seed=1,
                    device=dev, sort_every=8)
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    sim.x0.uniform_(0.0, 1.0, generator=gen).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
    sim.u0.normal_(0.0, 1.0, generator=gen)
    sim.u0[:sim.n_split].mul_(float(np.sqrt(kT / ME))); sim.u0[sim.n_split:].mul_(float(np.sqrt(kT / MP)))
    for _ in range(3):
        sim.step()
    tb = torch.zeros(2 * 148, dtype=torch.int64, device=dev)
    res = {}
    for stale in (0, 7):
        # fresh sort, then `stale` more steps without sorting
        sim.t = 0
        sim.step()
        for _ in range(stale):
            sim.step()
        sim.reinject()
        sim.Es.copy_(sim.E0)
        P = C.byref(sim.params)
        st = D.stream()
        sim.active.fill_(1)
        for first in (1, 0):
            times = []
            for rep in range(6):
                sim.acc.zero_()
                if rep == 5:
                    _lib.call("pic_dev_debug_cta_timer", D.ptr(tb))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.call("pic_dev_dd_picard_iter", P, D.ptr(sim.x0), D.ptr(sim.u0), D.ptr(sim.x1), D.ptr(sim.u1),
                          D.ptr(sim.active), D.ptr(sim.Es), D.ptr(sim.acc), first, D.ptr(sim.range_err), st)
                e1.record(); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            _lib.call("pic_dev_debug_cta_timer", None)
            t = tb.cpu().numpy().reshape(-1, 2)
            dur = (t[:, 1] - t[:, 0]) / 1e3
            end = (t[:, 1] - t[:, 0].min()) / 1e3
            res["stale%d_first%d" % (stale, first)] = dict(ms=[round(x, 4) for x in times], cta_us_min=float(dur.min()),
                                                           cta_us_med=float(np.median(dur)), cta_us_max=float(dur.max()),
                                                           slowest=[int(i) for i in np.argsort(-end)[:6]],
                                                           slowest_end_us=[float(x) for x in np.sort(-end)[:6] * -1])
    out[var] = res
    del sim
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
