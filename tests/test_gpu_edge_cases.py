"""Edge cases of the hot path through the public host mirror: empty and tiny stores, minimum
grids, odd particle counts, stores smaller than one kernel chunk, shards that hold one species
only -- every one against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import np_oracle as O


def relmax(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _sheath_inputs(N, Ng, seed):
    rs = np.random.RandomState(seed)
    dx, dt = 1e-5, 1e-12
    L = dx * (Ng - 1)
    h = N // 2
    x0 = rs.uniform(0, L, N)
    m = np.concatenate([np.full(h, O.me), np.full(N - h, O.mp)])
    q = np.concatenate([np.full(h, -O.e), np.full(N - h, O.e)])
    u0 = rs.normal(0, 1, N) * np.sqrt(O.kb * 116000. / m)
    E0 = rs.normal(0, 1e4, Ng)
    return dx, dt, L, x0, u0, q, m, E0


@pytest.mark.parametrize("N,Ng", [(0, 51), (1, 51), (2, 3), (3, 3), (7, 4), (16383, 9), (16385, 33), (32769, 51)])
def test_sheath_step_tiny_odd_and_chunk_boundary_sizes(N, Ng):
    from pypic_b200.sheath import SheathSim
    dx, dt, L, x0, u0, q, m, E0 = _sheath_inputs(N, Ng, 100 + N % 97)
    p2c = 1e10
    sim = SheathSim(N, Ng, dx, dt, p2c, kBT=(1.6e-18, 1.6e-18), carry_vw=False)
    sim.upload(x0, u0, E0=E0)
    k, r = sim.picard(); sim.check()
    act = np.ones(N)
    x1, u1, _, _, E1, j1, ko, ro, _ = O.dd_picard_step(x0, u0, np.zeros(N), np.zeros(N), q, m, act, E0, p2c, Ng, dx, dt, L, 1e-5, 20)
    out = sim.download()
    assert k == ko
    assert np.array_equal(out["active"][:N], act)
    if N:
        assert relmax(out["x0"][:N], x1) < 1e-12
    assert relmax(out["E0"], E1) < 1e-11


def test_sheath_single_species_shards():
    """A shard that holds only electrons or only ions (what rank 0 / rank 1 of a 2-rank run see)."""
    from pypic_b200 import _lib, device as D
    from pypic_b200.sheath import SheathSim
    import ctypes as C
    N, Ng = 40000, 65
    dx, dt, L, x0, u0, q, m, E0 = _sheath_inputs(N, Ng, 5)
    p2c = 1e10
    full = SheathSim(N, Ng, dx, dt, p2c, kBT=(1.6e-18, 1.6e-18), carry_vw=False, elide_u=False)
    full.upload(x0, u0, E0=E0); full.Es.copy_(full.E0); full.acc.zero_()
    _lib.call("pic_dev_dd_picard_iter", C.byref(full.params), D.ptr(full.x0), D.ptr(full.u0), D.ptr(full.x1), D.ptr(full.u1),
              D.ptr(full.active), D.ptr(full.Es), D.ptr(full.acc), 1, D.ptr(full.range_err), D.stream())
    ref_acc = full.acc.cpu().numpy().copy(); ref_x1 = full.x1.cpu().numpy().copy()
    tot = np.zeros_like(ref_acc)
    h = N // 2
    for lo, hi, ns in ((0, h, h), (h, N, 0)):
        s = SheathSim(hi - lo, Ng, dx, dt, p2c, n_split=ns, kBT=(1.6e-18, 1.6e-18), carry_vw=False, elide_u=False)
        s.upload(x0[lo:hi], u0[lo:hi], E0=E0); s.Es.copy_(s.E0); s.acc.zero_()
        _lib.call("pic_dev_dd_picard_iter", C.byref(s.params), D.ptr(s.x0), D.ptr(s.u0), D.ptr(s.x1), D.ptr(s.u1),
                  D.ptr(s.active), D.ptr(s.Es), D.ptr(s.acc), 1, D.ptr(s.range_err), D.stream())
        s.check()
        assert np.array_equal(s.x1.cpu().numpy(), ref_x1[lo:hi])
        tot += s.acc.cpu().numpy()
    assert relmax(tot[:2 * Ng], ref_acc[:2 * Ng]) < 1e-13 and np.array_equal(tot[2 * Ng:], ref_acc[2 * Ng:])


@pytest.mark.parametrize("N,Ng", [(0, 16), (1, 16), (5, 2), (16385, 8), (20001, 7)])
def test_periodic_sims_tiny_and_odd_sizes(N, Ng):
    from pypic_b200.periodic import ExplicitSim, PeriodicImplicitSim
    rs = np.random.RandomState(N + Ng)
    dx = 1e-5; dt = 1e-12; L = dx * Ng
    x0 = rs.uniform(0, L, N); v0 = rs.normal(0, 1e6, N); E0 = rs.normal(0, 1e4, Ng)
    sim = PeriodicImplicitSim(N, Ng, dx, dt, L, 1e9, tol=1e-30, maxiter=1)
    sim.upload(x0, v0, E0)
    k, r = sim.push(); sim.check()
    q = -np.ones(N) * O.e; m = np.ones(N) * O.me
    x1, v1, E1, j1, ko, ro = O.pypic_particle_push_p(x0, v0, q, m, E0, np.zeros(Ng), N, Ng, 1e9, dx, dt, L, 1e-30, 1)
    out = sim.download()
    assert np.array_equal(out["x0"], x1) and np.array_equal(out["v0"], v1)
    assert relmax(out["E0"], E1) < 1e-12
    # explicit: one step keeps every particle inside [0, L+dx)
    Le = dx * (Ng - 1)
    xe = rs.uniform(0, Le + dx, N)
    ex = ExplicitSim(N, Ng, dx, dt, 1e9)
    ex.upload(xe, v0)
    ex.step(); ex.check()
    o = ex.download()
    assert len(o["x"]) == N and (N == 0 or (o["x"].min() >= 0 and o["x"].max() < Le + dx))
    assert np.isfinite(o["E"]).all()


def test_gc_store_empty_and_single_particle():
    from pypic_b200.gcstore import GridDev, ParticleStore
    grid = GridDev(16, 1e-3, 6e5)
    for N in (0, 1):
        r = np.zeros((N, 7)); r[:, 0] = 5e-4; r[:, 3] = 1e4
        st = ParticleStore.from_arrays(r, 1.0, O.mp, 1e9, Z=1, B=(0.1, 2.0, 0.0))
        grid.weight_particles_to_grid_boltzmann(st, 1e-10)
        hits = st.push_6D(1e-10, grid)
        st.check(); grid.check()
        assert hits == 0 and st.N == N
        if N:
            ref = O.gc_push_6D(r, np.zeros(1), np.array([0.1, 2.0, 0.0]), 1.0, O.mp, 1e-10)
            assert np.array_equal(st.r_host(), ref)
