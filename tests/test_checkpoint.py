"""Checkpoint / restart: file format on CPU, bit-identical continuation on GPU."""
import numpy as np
import pytest

from oracle import np_oracle as O


def test_rng_state_round_trip_cpu(tmp_path):
    """The legacy MT19937 state (incl. the cached second gaussian) survives the .npz."""
    from pypic_b200 import checkpoint as K
    np.random.seed(7)
    np.random.normal(size=3)                      # leaves a cached gaussian behind
    st = K._rng_state()
    np.savez(tmp_path / "s.npz", **st)
    a = [np.random.uniform(), np.random.normal(), np.random.normal()]
    np.random.seed(99)
    K._set_rng_state(np.load(tmp_path / "s.npz"))
    b = [np.random.uniform(), np.random.normal(), np.random.normal()]
    assert a == b


def test_reference_particle_records_cpu():
    from pypic_b200 import checkpoint as K
    N = 5
    state = dict(N=N, r=np.arange(N * 7, dtype=float).reshape(N, 7), m=np.full(N, O.mp), charge_state=np.ones(N),
                 p2c=np.full(N, 2.0), Z=np.ones(N, dtype=np.int32), active=np.ones(N, dtype=np.int8),
                 at_wall=np.zeros(N, dtype=np.int8), from_wall=np.zeros(N, dtype=np.int8))
    recs = K.to_reference_particles(state)
    assert len(recs) == N and recs[3]["r"][0] == 21.0 and recs[0]["Z"] == 1 and recs[2]["active"] == 1


@pytest.mark.gpu
def test_sheath_restart_continues_bit_identically(tmp_path):
    """6 steps straight == 3 steps, save, fresh simulation, load, 3 steps (host MT19937 draws)."""
    from pypic_b200 import checkpoint as K
    from pypic_b200.sheath import SheathSim
    N, Ng = 30000, 65
    dx, dt = 1e-5, 1e-12
    L = dx * (Ng - 1)
    kT = O.kb * 116000.
    rs = np.random.RandomState(2)
    x0 = rs.uniform(0, L, N)
    u0 = np.concatenate([rs.normal(0, np.sqrt(kT / O.me), N // 2), rs.normal(0, np.sqrt(kT / O.mp), N - N // 2)])
    v0 = rs.normal(size=N); w0 = rs.normal(size=N)

    def fresh():
        s = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=True, rng="host")
        return s

    def run(s, n):
        its = []
        for _ in range(n):
            its.append(s.step()[0])
        return its
    np.random.seed(1)
    a = fresh(); a.upload(x0, u0, v0, w0)
    its_a = run(a, 6)
    np.random.seed(1)
    b = fresh(); b.upload(x0, u0, v0, w0)
    its_b = run(b, 3)
    K.save_sheath(b, tmp_path / "ck.npz")
    np.random.seed(12345)                         # scramble the stream: load must restore it
    c = fresh()
    K.load_sheath(c, tmp_path / "ck.npz")
    its_b += run(c, 3)
    assert its_a == its_b
    oa, oc = a.download(), c.download()
    # deposits are atomically accumulated (order-dependent round-off) -> fields to 1e-12,
    # flags and the injected particles identical
    assert np.array_equal(oa["active"], oc["active"])
    for k in ("x0", "u0", "v0", "w0"):
        assert np.max(np.abs(oa[k] - oc[k])) <= 1e-12 * np.max(np.abs(oa[k]))
    assert np.max(np.abs(oa["E0"] - oc["E0"])) <= 1e-11 * np.max(np.abs(oa["E0"]))


@pytest.mark.gpu
def test_gc_store_and_grid_round_trip(tmp_path):
    from pypic_b200 import checkpoint as K
    from pypic_b200.gcstore import GridDev, ParticleStore
    rs = np.random.RandomState(3)
    N, ng, Lg = 5000, 64, 1e-3
    r = rs.normal(size=(N, 7)); r[:, 0] = rs.uniform(0, Lg, N)
    st = ParticleStore.from_arrays(r, 1.0, O.mp, 3e9, Z=1, active=(rs.uniform(size=N) < 0.9).astype(np.int8), B=(0.1, 2.0, 0))
    grid = GridDev(ng, Lg, 6e5)
    grid.weight_particles_to_grid_boltzmann(st, 1e-10)
    grid.solve_for_phi_dirichlet_boltzmann(); grid.differentiate_phi_to_E_dirichlet()
    grid.add_particles(3e9)
    K.save_gc(st, grid, tmp_path / "gc.npz")
    st2, g2 = K.load_gc(tmp_path / "gc.npz")
    assert st2.N == N and np.array_equal(st2.r_host(), st.r_host())
    for k in ("active", "at_wall", "from_wall"):
        assert np.array_equal(st2.flags_host()[k], st.flags_host()[k])
    for name in ("rho", "phi", "E", "n", "state"):
        assert np.array_equal(getattr(g2, name).cpu().numpy(), getattr(grid, name).cpu().numpy())
    assert g2.added_particles == grid.added_particles and g2.n0 == grid.n0
    # both continue identically
    h1 = st.push_6D(1e-10, grid); h2 = st2.push_6D(1e-10, g2)
    assert h1 == h2 and np.array_equal(st2.r_host(), st.r_host())
