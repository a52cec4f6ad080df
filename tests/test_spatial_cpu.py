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
