import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_oracle as O
from pypic_b200 import _lib, device as D
from pypic_b200.sheath import SheathSim
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_sheath import _one_iter_inputs
N, Ng = 200000, 257
dx, L, dt, x0, u0, q, m, E0 = _one_iter_inputs(N, Ng, 5)
p2c = 1e9
qm = q / m
Ei = O.dd_interpolateField(E0, x0, Ng, dx)
x1 = x0 + dt * u0 + dt * dt * qm * Ei * 0.5
u1 = u0 + dt * qm * Ei
qm_e = qm[0]
x1_e = x0 + dt * u0 + dt * dt * qm_e * Ei * 0.5
nbad = 0
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for trial in range(trials):
    b = SheathSim(N, Ng, dx, dt, p2c, kBT=(1.6e-18, 1.6e-18), carry_vw=False, rng="philox", sort_every=0)
    b.upload(x0, u0, E0=E0)
    b.Es.copy_(b.E0)
    P = C.byref(b.params)
    for rep in range(3):
        b.acc.zero_(); b.x1.fill_(-7.0); b.u1.fill_(-7.0)
        _lib.call("pic_dev_dd_picard_iter", P, D.ptr(b.x0), D.ptr(b.u0), D.ptr(b.x1), D.ptr(b.u1), D.ptr(b.active),
                  D.ptr(b.Es), D.ptr(b.acc), 1, D.ptr(b.range_err), D.stream())
        gx1 = b.x1.cpu().numpy(); gu1 = b.u1.cpu().numpy()
        w = np.where((gx1 != x1) | (gu1 != u1))[0]
        if len(w):
            nbad += 1
            print("trial", trial, "rep", rep, "mismatches", len(w), "first", w[:4], "last", w[-1])
            for i in w[:4]:
                ch, r = divmod(int(i), 16384); wp, r2 = divmod(r, 1024); row, r3 = divmod(r2, 64)
                ui = (gx1[i] - x0[i]) / dt
                j = int(np.argmin(np.abs(u0 - ui)))
                print("   i", i, "chunk", ch, "warp", wp, "row", row, "lane", r3 // 2, "gx1", gx1[i], "x1", x1[i], "x1 with e- const", x1_e[i],
                      "implied u", ui, "u0", u0[i], "nearest u0 idx", j, u0[j], "gu1", gu1[i], "u1", u1[i])
    del b
print("bad launches:", nbad, "of", trials * 3)
