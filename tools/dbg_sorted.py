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
for sort_every in (0,):
    b = SheathSim(N, Ng, dx, dt, p2c, kBT=(1.6e-18, 1.6e-18), carry_vw=False, rng="philox", sort_every=sort_every)
    b.upload(x0, u0, E0=E0)
    if sort_every:
        b.sort_by_cell()
    X0 = b.x0.cpu().numpy().copy(); U0 = b.u0.cpu().numpy().copy()
    b.Es.copy_(b.E0)
    P = C.byref(b.params)
    act = np.ones(N); qm = q / m
    xh_prev = None
    for it in range(3):
        b.acc.zero_()
        _lib.call("pic_dev_dd_picard_iter", P, D.ptr(b.x0), D.ptr(b.u0), D.ptr(b.x1), D.ptr(b.u1), D.ptr(b.active),
                  D.ptr(b.Es), D.ptr(b.acc), 1 if it == 0 else 0, D.ptr(b.range_err), D.stream())
        a = act == 1
        xs = X0 if it == 0 else xh_prev
        Ei = O.dd_interpolateField(E0, xs[a], Ng, dx)
        x1 = np.zeros(N); u1 = np.zeros(N); xh = np.zeros(N); uh = np.zeros(N)
        x1[a] = X0[a] + dt * U0[a] + dt * dt * qm[a] * Ei * 0.5
        u1[a] = U0[a] + dt * qm[a] * Ei
        xh[a] = (X0[a] + x1[a]) * 0.5; uh[a] = (U0[a] + u1[a]) * 0.5
        right = a & ((X0 >= L) | (xh >= L) | (x1 >= L)); act[right] = 0
        left = (act == 1) & ((X0 <= 0) | (xh <= 0) | (x1 <= 0)); act[left] = -1
        jh = np.zeros(Ng); j1 = np.zeros(Ng); s = act == 1
        ih, wL, wR = O.dd_index_weights(xh[s], dx)
        np.add.at(jh, ih, q[s] * uh[s] * p2c * wL / dx); np.add.at(jh, ih + 1, q[s] * uh[s] * p2c * wR / dx)
        i1, wL1, wR1 = O.dd_index_weights(x1[s], dx)
        np.add.at(j1, i1, q[s] * u1[s] * p2c * wL1 / dx); np.add.at(j1, i1 + 1, q[s] * u1[s] * p2c * wR1 / dx)
        g = b.acc.cpu().numpy()
        gx1 = b.x1.cpu().numpy(); gu1 = b.u1.cpu().numpy(); gact = b.active.cpu().numpy()
        ejh = np.abs(g[:Ng] - jh) / np.abs(jh).max(); ej1 = np.abs(g[Ng:2 * Ng] - j1) / np.abs(j1).max()
        print("sort", sort_every, "it", it, "x1 eq", np.array_equal(gx1, x1), "u1 eq", np.array_equal(gu1, u1), "act eq",
              np.array_equal(gact, act), "jh err", ejh.max(), int(ejh.argmax()), "j1 err", ej1.max(), int(ej1.argmax()),
              "counts", g[2 * Ng:], (left[:N // 2].sum(), left[N // 2:].sum(), right[:N // 2].sum(), right[N // 2:].sum()))
        if ejh.max() > 1e-10:
            bad = np.where(ejh > 1e-10)[0]; print("  bad jh nodes", bad[:20], len(bad))
        if not np.array_equal(gx1, x1):
            w = np.where(gx1 != x1)[0]; print("  x1 differs at", w[:10], len(w), gx1[w[:3]], x1[w[:3]])
            for i in w[:16]:
                ui = (gx1[i] - X0[i]) / dt
                j = int(np.argmin(np.abs(U0 - ui)))
                du = (gu1[i] - u1[i])
                print("   i", i, "U0", U0[i], "implied u", ui, "nearest U0 idx", j, U0[j], "gu1-u1", du, "gu1", gu1[i], "u1", u1[i])
        xh_prev = xh
