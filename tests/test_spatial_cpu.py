"""Host logic of the slab decomposition (pypic_b200/spatial.py) that needs no GPU."""
import numpy as np

from pypic_b200.spatial import slab_bounds, swap_remove_plan


def test_slab_bounds_cover_the_grid():
    for Ng in (51, 4097, 1000001):
        for W in (1, 2, 3, 8):
            b = slab_bounds(Ng, W)
            assert b[0] == 0 and b[-1] == Ng - 1 and len(b) == W + 1
            assert all(y > x for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1


def _apply(n, holes, arrivals):
    """Simulates the plan on labelled slots; returns the surviving labels."""
    slots = {i: ("p", i) for i in range(n)}
    adst, msrc, mdst, new_n = swap_remove_plan(n, holes, len(arrivals))
    for s, d in zip(msrc, mdst):
        slots[int(d)] = slots[int(s)]
    for a, d in zip(arrivals, adst):
        slots[int(d)] = ("a", a)
    return [slots[i] for i in range(new_n)], new_n


def test_swap_remove_plan_keeps_every_survivor_once():
    rs = np.random.RandomState(0)
    for n, nh, na in ((10, 0, 0), (10, 3, 0), (10, 0, 4), (10, 3, 3), (10, 5, 2), (10, 2, 6), (200, 57, 13), (200, 13, 57),
                      (8, 8, 0), (8, 8, 3), (5, 2, 0)):
        for _ in range(5):
            holes = np.sort(rs.choice(n, nh, replace=False)) if nh else np.zeros(0, np.int64)
            arrivals = list(range(1000, 1000 + na))
            out, new_n = _apply(n, holes, arrivals)
            assert new_n == n - nh + na and len(out) == new_n
            survivors = sorted(i for i in range(n) if i not in set(holes.tolist()))
            assert sorted(v for k, v in out if k == "p") == survivors
            assert sorted(v for k, v in out if k == "a") == arrivals


def test_exchange_plan_reproduces_the_global_sum():
    """Emulates W ranks on the CPU: every rank deposits only on its own nodes and on `G` guard nodes
    either side; pack -> all-gather -> unpack must give, on EVERY rank, the sum over ranks of the
    accumulators (jh, j1) and of the 4 absorbed counts."""
    from pypic_b200.spatial import exchange_plan
    rs = np.random.RandomState(1)
    for Ng, W, G in ((257, 2, 16), (513, 3, 8), (4097, 8, 16), (1000, 5, 3)):
        cb = slab_bounds(Ng, W)
        accs = []
        for r in range(W):
            acc = np.zeros(2 * Ng + 5)
            lo, hi = max(cb[r] - G, 0), min(cb[r + 1] + G + 1, Ng)          # nodes this rank can touch
            acc[lo:hi] = rs.normal(size=hi - lo); acc[Ng + lo:Ng + hi] = rs.normal(size=hi - lo)
            acc[2 * Ng:2 * Ng + 4] = rs.randint(0, 50, 4)
            accs.append(acc)
        want = np.sum(accs, 0)
        plans = [exchange_plan(Ng, G, cb, r) for r in range(W)]
        gathered = np.concatenate([accs[r][plans[r]["pack"]] for r in range(W)])     # what all_gather_into_tensor delivers
        for r in range(W):
            pl = plans[r]
            out = accs[r].copy()
            out[:2 * Ng] = gathered[pl["unpack"]]
            np.add.at(out, pl["add_dst"], gathered[pl["add_src"]])
            out[2 * Ng:2 * Ng + 4] = gathered[pl["counts"]].reshape(W, 4).sum(0)
            assert np.allclose(out[:2 * Ng + 4], want[:2 * Ng + 4], rtol=0, atol=1e-12), (Ng, W, G, r)
            assert out[2 * Ng + 4] == 0.0


def _slab_pack_model(raw, Ng, c0, c1, G, rank, W):
    """NumPy restatement of slab_pack_k (csrc/slab_kernels.cu): the message of one rank."""
    B = 2 * G + 1
    msg = np.zeros(4 * B + 10)
    jh, j1 = raw[:Ng], raw[Ng:2 * Ng]
    if rank > 0:
        msg[0:B] = jh[c0 - G:c0 + G + 1]; msg[B:2 * B] = j1[c0 - G:c0 + G + 1]
    if rank < W - 1:
        msg[2 * B:3 * B] = jh[c1 - G:c1 + G + 1]; msg[3 * B:4 * B] = j1[c1 - G:c1 + G + 1]
    b0, b1 = max(c0 - G, 0), min(c1 + G + 1, Ng)
    sh, s1 = jh[b0:b1].sum(), j1[b0:b1].sum()
    if rank == 0:
        sh += jh[1]; s1 += j1[1]; msg[4 * B + 6:4 * B + 8] = jh[1], j1[1]
    if rank == W - 1:
        sh += jh[Ng - 2]; s1 += j1[Ng - 2]; msg[4 * B + 8:4 * B + 10] = jh[Ng - 2], j1[Ng - 2]
    msg[4 * B], msg[4 * B + 1] = sh, s1
    msg[4 * B + 2:4 * B + 6] = raw[2 * Ng:2 * Ng + 4]
    return msg


def _slab_field_model(raw, gath, wall_cum, E0, Es, Ng, c0, c1, G, rank, W, dx, dt, p2c, q):
    """NumPy restatement of slab_field_k: (E1, Eh, j1) on the band, residual^2 and field energy over the owned nodes."""
    eps0 = 8.854E-12
    B = 2 * G + 1; M = 4 * B + 10
    msgs = gath.reshape(W, M)
    w = wall_cum + msgs[:, 4 * B + 2:4 * B + 6].sum(0)
    wallL = w[0] * (dx * q[0] * p2c / dt) + w[1] * (dx * q[1] * p2c / dt)
    wallR = w[2] * (-dx * q[0] * p2c / dt) + w[3] * (-dx * q[1] * p2c / dt)
    meanh = ((msgs[:, 4 * B].sum() + wallL) + wallR) / Ng
    b0, b1 = max(c0 - G, 0), min(c1 + G + 1, Ng)
    a = raw[:Ng].copy(); b = raw[Ng:2 * Ng].copy()
    if rank > 0:
        a[c0 - G:c0 + G + 1] += msgs[rank - 1, 2 * B:3 * B]; b[c0 - G:c0 + G + 1] += msgs[rank - 1, 3 * B:4 * B]
    if rank < W - 1:
        a[c1 - G:c1 + G + 1] += msgs[rank + 1, 0:B]; b[c1 - G:c1 + G + 1] += msgs[rank + 1, B:2 * B]
    mine = msgs[rank, 4 * B + 6:]
    if b0 == 0:
        a[0] = (a[0] + wallL) + mine[0]; b[0] = (b[0] + wallL) + mine[1]
    if b1 == Ng:
        a[Ng - 1] = (a[Ng - 1] + wallR) + mine[2]; b[Ng - 1] = (b[Ng - 1] + wallR) + mine[3]
    band = slice(b0, b1)
    e1 = E0[band] + (dt / eps0) * (meanh - a[band])
    eh = (e1 + E0[band]) * 0.5
    o0, o1 = c0, (c1 + 1 if rank == W - 1 else c1)
    own = slice(o0 - b0, o1 - b0)
    d = Es[band] - eh
    return band, e1, eh, b[band], float((d[own] ** 2).sum()), float((eps0 * e1[own] ** 2 * dx / 2.).sum())


def test_distributed_field_update_scheme_reproduces_the_global_update():
    """The message scheme of the slab decomposition's distributed field update (two boundary bands, the sum of all
    raw deposits, four counts and the fold nodes per rank) against the reference's global formulas
    (PIC_L_DD.py:55-66, 516-527; oracle.np_oracle.dd_picard_step) with W ranks emulated in NumPy."""
    rs = np.random.RandomState(2)
    eps0 = 8.854E-12
    dx, dt, p2c, q = 1e-5, 1e-12, 3.7e9, (-1.602e-19, 1.602e-19)
    for Ng, W, G in ((257, 1, 16), (257, 2, 16), (513, 3, 8), (4097, 8, 16), (1000, 5, 3)):
        cb = slab_bounds(Ng, W)
        raws = []
        for r in range(W):
            raw = np.zeros(2 * Ng + 4)
            lo, hi = max(cb[r] - G, 0), min(cb[r + 1] + G + 1, Ng)
            raw[lo:hi] = rs.normal(size=hi - lo) * 1e3; raw[Ng + lo:Ng + hi] = rs.normal(size=hi - lo) * 1e3
            if r == 0:
                raw[2 * Ng:2 * Ng + 2] = rs.randint(0, 50, 2)
            if r == W - 1:
                raw[2 * Ng + 2:2 * Ng + 4] = rs.randint(0, 50, 2)
            raws.append(raw)
        wall_cum = rs.randint(0, 20, 4).astype(np.float64)
        E0 = rs.normal(0, 1e4, Ng); Es = E0 + rs.normal(0, 1.0, Ng)
        # ---- the global update
        tot = np.sum(raws, 0)
        w = wall_cum + tot[2 * Ng:]
        wallL = w[0] * (dx * q[0] * p2c / dt) + w[1] * (dx * q[1] * p2c / dt)
        wallR = w[2] * (-dx * q[0] * p2c / dt) + w[3] * (-dx * q[1] * p2c / dt)
        jh, j1 = tot[:Ng].copy(), tot[Ng:2 * Ng].copy()
        jh[0] += wallL; jh[-1] += wallR; j1[0] += wallL; j1[-1] += wallR
        jh[0] += tot[1]; jh[-1] += tot[Ng - 2]; j1[0] += tot[Ng + 1]; j1[-1] += tot[2 * Ng - 2]
        E1 = E0 + (dt / eps0) * (np.average(jh) - jh)
        Eh = (E1 + E0) * 0.5
        # ---- the slabs
        gath = np.concatenate([_slab_pack_model(raws[r], Ng, cb[r], cb[r + 1], G, r, W) for r in range(W)])
        rr = ee = 0.0
        scale = np.max(np.abs(E1))
        prev = None
        for r in range(W):
            band, e1, eh, b, rr_r, ee_r = _slab_field_model(raws[r], gath, wall_cum, E0, Es, Ng, cb[r], cb[r + 1], G, r, W,
                                                            dx, dt, p2c, q)
            assert np.max(np.abs(e1 - E1[band])) <= 1e-13 * scale and np.max(np.abs(eh - Eh[band])) <= 1e-13 * scale
            assert np.max(np.abs(b - j1[band])) <= 1e-12 * np.max(np.abs(j1))
            if prev is not None:                 # the nodes two neighbours both compute carry identical bits
                pb, pe1 = prev
                lo = band.start
                assert np.array_equal(pe1[lo - pb.start:], e1[:pb.stop - lo])
            prev = (band, e1)
            rr += rr_r; ee += ee_r
        assert abs(rr - ((Es - Eh) ** 2).sum()) <= 1e-10 * rr
        assert abs(ee - (eps0 * E1 ** 2 * dx / 2.).sum()) <= 1e-12 * ee
