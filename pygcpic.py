"""Drop-in replacement for the hot path of the reference's pygcpic.py: the ``Particle`` and
``Grid`` classes (same constructor arguments, attributes and method names), the particle
source generators, and a device-resident time loop.  Every numerical method -- gather, Boris
push, 6D<->guiding-centre transforms, GC RK4, boundary tests, CIC deposit with the
Boltzmann-electron reference density, smoothing, the linear and Newton-Boltzmann Poisson
solves, E = -grad(phi) -- executes as a hand-written sm_100a CUDA kernel of libpic_b200.so
(through pypic_b200).  There is no CPU fallback.

Two ways to use it:

* ``Particle`` / ``Grid`` objects exactly like the reference (one object per particle, NumPy
  attributes).  Each numerical method call stages its operands to the GPU and back, so the
  reference's doctests and object-level drivers run unchanged.  ``Grid`` methods that take a
  list of particles stage the whole list as structure-of-arrays in ONE upload.
* ``ParticleStore`` / ``GridDev`` (re-exported from pypic_b200.gcstore) keep the particles
  resident in HBM as structure-of-arrays; ``run_sheath`` is the per-timestep loop of
  pic_bca_aps (pygcpic.py:1486-1563) on that store.

Monte-Carlo ionisation (``Particle.attempt_first_ionization`` / ``attempt_nth_ionization`` and
``run_sheath(..., ionize_Te=)``) and the IEAD wall-hit histograms (``pypic_b200.ops`` /
``pic_dev_gc_iead_hist``) are built.  Out of scope: the F-TRIDYN (BCA)
binary itself and the plotting drivers ``pic_iead`` / ``pic_bca_aps`` / ``pic_bca`` that wrap it.
"""
import numpy as np
import torch

from pypic_b200 import ops, device as _D
from pypic_b200.gcstore import GridDev, ParticleStore  # noqa: F401  (device-resident API)

# physical constants (pygcpic.py:13-17)
epsilon0 = 8.854e-12
e = 1.602e-19
mp = 1.67e-27
me = 9.11e-31
kb = 1.38e-23


def gaussian_distribution(x, mu, sigma):
    """pygcpic.py:31-32."""
    return 1. / np.sqrt(2. * np.pi * sigma * sigma) * np.exp(-(x - mu)**2 / (2. * sigma**2))


def weighted_gaussian(x, mu, sigma):
    """pygcpic.py:757-758."""
    return gaussian_distribution(x, mu, sigma) * np.abs(x)


def _store_of(particles, B=None):
    """Stages a list of Particle objects as a structure-of-arrays ParticleStore (one upload)."""
    n = len(particles)
    r = np.empty((n, 7))
    cs = np.empty(n); m = np.empty(n); p2c = np.empty(n); act = np.empty(n, dtype=np.int8)
    for i, p in enumerate(particles):
        r[i] = p.r; cs[i] = p.charge_state; m[i] = p.m; p2c[i] = p.p2c; act[i] = p.active
    if B is None:
        B = particles[0].B if n else np.zeros(3)
    return ParticleStore.from_arrays(r, cs, m, p2c, active=act, B=tuple(float(b) for b in B))


class Particle:
    """pygcpic.Particle (pygcpic.py:70-720): host-side object, device-side arithmetic."""

    def __init__(self, m, charge_state, p2c, T, Z, B0=np.zeros(3), E0=np.zeros(3), grid=None, vx=0.):
        self.r = np.zeros(7)
        self.charge_state = charge_state
        self.Z = Z
        self.m = m
        self.T = T
        self.p2c = p2c
        self.vth = np.sqrt(kb * self.T / self.m)
        self.mode = 0          # 0: 6D, 1: guiding centre
        self.E = E0            # NOTE: like the reference, the default array is shared by all instances
        self.B = B0
        self.active = 1
        self.at_wall = 0
        self.from_wall = 0
        if grid is not None:
            self._initialize_6D(grid, vx=vx)

    def __repr__(self):
        return f'Particle({self.m}, {self.charge_state}, {self.p2c}, {self.T}, {self.Z})'

    def is_active(self):
        return self.active == 1

    # -- trivial accessors (host) --------------------------------------------------------
    @property
    def speed(self):
        return np.sqrt(self.r[3]**2 + self.r[4]**2 + self.r[5]**2)

    @speed.setter
    def speed(self, speed):
        u = self.v / np.linalg.norm(self.v)
        self.v = u * speed

    @property
    def x(self):
        return self.r[0]

    @x.setter
    def x(self, x0):
        self.r[0] = x0

    @property
    def y(self):
        return self.r[1]

    @property
    def z(self):
        return self.r[2]

    @property
    def v_x(self):
        return self.r[3]

    @v_x.setter
    def v_x(self, v_x):
        self.r[3] = v_x

    @property
    def v(self):
        return self.r[3:6]

    @v.setter
    def v(self, v0):
        self.r[3:6] = v0

    def get_angle_wrt_wall(self, use_degrees=True):
        """pygcpic.py:228-259."""
        v = self.r[3:6]
        alpha = np.arctan2(np.sqrt(v[1]**2 + v[2]**2), np.abs(v[0]))
        return alpha * 180. / np.pi if use_degrees else alpha

    @property
    def kinetic_energy(self):
        return 0.5 * self.m * self.speed**2

    def _initialize_6D(self, grid, vx=0.):
        """pygcpic.py:277-304: draw order uniform(x) then normal x3 from the global legacy stream."""
        self.r[0] = np.random.uniform(0.0, grid.length)
        self.r[1:3] = 0.0
        self.r[3:6] = np.random.normal(0.0, self.vth, 3) + vx
        self.r[6] = 0.0

    def set_x_direction(self, direction):
        """pygcpic.py:306-323."""
        if not isinstance(direction, str):
            raise TypeError('particle.set_x_direction(direction) received a non-string type for direction')
        if direction.lower() == 'left':
            self.r[3] = -abs(self.r[3])
        elif direction.lower() == 'right':
            self.r[3] = abs(self.r[3])
        else:
            raise ValueError('particle.set_x_direction() received neither right nor left')

    # -- device arithmetic -----------------------------------------------------------------
    def _dev(self):
        return ParticleStore.from_arrays(self.r[None, :], self.charge_state, self.m, self.p2c, active=[1],
                                         B=tuple(float(b) for b in self.B), Eyz=(float(self.E[1]), float(self.E[2])))

    def interpolate_electric_field_dirichlet(self, grid):
        """pygcpic.py:325-348 (mirrored weights; writes E[0] of the -- possibly shared -- E array)."""
        self.E[0] = ops.gc_interpolate(grid.E, self.r[0], grid.ng, grid.dx)[0]

    def push_6D(self, dt):
        """pygcpic.py:460-507: Boris push in the particle's E, B."""
        s = self._dev()
        s.push_6D_given_E(dt, _D.to_dev(np.array([float(self.E[0])]), s.dev))
        self.r[:] = s.r_host()[0]

    def transform_6D_to_GC(self):
        """pygcpic.py:509-551."""
        s = self._dev()
        s.transform_6D_to_GC()
        self.r[:] = s.r_host()[0]
        self.mode = 1

    def transform_GC_to_6D(self):
        """pygcpic.py:553-596: the gyro-phase vector a = uniform(0,1,3) comes from the global
        legacy stream, like the reference."""
        a = np.random.uniform(0.0, 1.0, 3)
        s = self._dev()
        s.transform_GC_to_6D(a[None, :])
        self.r[:] = s.r_host()[0]
        self.mode = 0

    def push_GC(self, dt):
        """pygcpic.py:598-645: RK4 of the guiding-centre equations, E and B frozen over the step."""
        s = self._dev()
        s.push_GC_given_E(dt, _D.to_dev(np.array([float(self.E[0])]), s.dev))
        self.r[:] = s.r_host()[0]

    def apply_BCs_periodic(self, grid):
        """pygcpic.py:647-666."""
        from pypic_b200 import _lib
        t = _D.to_dev(np.array([float(self.r[0])]), _D.require_cuda())
        _lib.call("pic_dev_wrap_periodic", _D.ptr(t), 1, float(grid.length), _D.stream())
        self.r[0] = t.cpu().numpy()[0]

    def apply_BCs_dirichlet(self, grid):
        """pygcpic.py:668-689."""
        from pypic_b200 import _lib
        dev = _D.require_cuda()
        x = _D.to_dev(np.array([float(self.r[0])]), dev)
        a = torch.tensor([int(self.active)], dtype=torch.int8, device=dev)
        w = torch.tensor([int(self.at_wall)], dtype=torch.int8, device=dev)
        _lib.call("pic_dev_gc_apply_bcs", _D.ptr(x), _D.ptr(a), _D.ptr(w), 1, float(grid.length), _D.stream())
        self.active = int(a.item()); self.at_wall = int(w.item())

    def reactivate(self, distribution, grid, time, p2c, m, charge_state, Z):
        """pygcpic.py:691-720."""
        self.r = next(distribution)
        self.p2c = p2c
        self.m = m
        self.charge_state = charge_state
        self.Z = Z
        self.r[6] = time
        self.active = 1
        self.at_wall = 0
        self.from_wall = 0
        grid.add_particles(p2c)

    def _ionization_attempt(self, table_state, dt, temperature, grid):
        """Common body of pygcpic.py:385-399 / 440-458: rate from the tabulated coefficients, plasma
        density gathered at the particle (plain CIC weights, on the device), probability
        density^2 * rate * dx * dt / p2c, ONE uniform from the global legacy stream, and -- in both
        routines -- the ionisation only if charge_state == 0 (tested after the draw)."""
        from pypic_b200 import ionization, ops
        if (int(self.Z), int(table_state)) not in ionization._TABLES:
            raise UnboundLocalError("no rate table for Z=%r, charge_state=%r (the reference's tables cover H and B 0..2+)"
                                    % (self.Z, table_state))
        ionization_rate = ionization.rate(self.Z, table_state, temperature)
        density = float(ops.dd_interpolate(grid.n, self.x, grid.ng, grid.dx)[0])
        probability = density**2 * ionization_rate * grid.dx * dt / self.p2c
        if np.random.uniform(0., 1.) < probability and self.charge_state == 0.:
            return True
        return False

    def attempt_first_ionization(self, dt, temperature, grid):
        """pygcpic.py:350-399 (the table is chosen by Z alone)."""
        if self._ionization_attempt(0, dt, temperature, grid):
            self.charge_state = 1
            grid.add_particles(self.p2c)

    def attempt_nth_ionization(self, dt, temperature, grid):
        """pygcpic.py:401-458 (boron only; the table follows the charge state)."""
        if int(self.Z) != 5:
            raise UnboundLocalError("attempt_nth_ionization has tables for boron (Z = 5) only, as in the reference")
        if self._ionization_attempt(self.charge_state, dt, temperature, grid):
            print(f'Ionized boron from {self.charge_state} to {self.charge_state+1}!')
            self.charge_state += 1
            grid.add_particles(self.p2c)


def source_distribution_6D(grid, Ti, mass, vx=0.):
    """pygcpic.py:723-755 (host generator on the global legacy stream)."""
    while True:
        vth = np.sqrt(kb * Ti / mass)
        r = np.empty(7)
        r[0] = np.random.normal(grid.length / 2, grid.length / 12.0)
        r[0] %= grid.length
        r[1:3] = 0.
        r[3:6] = np.random.normal(0.0, vth, 3) + vx
        yield r


def flux_distribution_6D(grid, Ti, mass, vx=0., gamma=0., vx_pert=0.):
    """pygcpic.py:760-778."""
    while True:
        vth = np.sqrt(kb * Ti / mass)
        r = np.empty(7)
        r[0] = grid.length - grid.dx * np.random.uniform(0., 1.)
        r[1:3] = 0.
        r[3:6] = np.random.normal(0.0, vth, 3)
        vels = np.linspace(-6 * vth, 6 * vth, 100)
        dist = np.array([weighted_gaussian(vel, vx, vth) for vel in vels])
        dist /= np.sum(dist)
        r[3] = -np.abs(np.random.choice(vels, p=dist)) + np.random.uniform(-1, 1) * (vels[1] - vels[0]) / 2.
        r[3] += vx
        if np.random.uniform(0, 1) < gamma:
            r[3] = vx_pert * vth
        yield r


class Grid:
    """pygcpic.Grid (pygcpic.py:780-1117) with NumPy attributes; every method runs on the GPU."""

    def __init__(self, ng, length, Te, bc='dirichlet-dirichlet'):
        self.ng = ng
        assert self.ng > 1, 'Number of grid points must be greater than 1'
        self.length = length
        assert self.length > 0.0, 'Length must be greater than 0'
        self.domain = np.linspace(0.0, length, ng)
        self.dx = self.domain[1] - self.domain[0]
        self.rho = np.zeros(ng)
        self.phi = np.zeros(ng)
        self.E = np.zeros(ng)
        self.n = np.zeros(ng)
        self.n0 = None
        self.rho0 = None
        self.Te = Te
        self.ve = np.sqrt(8. / np.pi * kb * self.Te / me)
        self.added_particles = 0
        self.bc = bc
        if bc == 'dirichlet-dirichlet':
            self._fill_laplacian_dirichlet()
        elif bc == 'dirichlet-neumann':
            self._fill_laplacian_dirichlet_neumann()
            print(self.A)
        elif not isinstance(bc, str):
            raise TypeError('bc must be a string')
        else:
            raise ValueError('Unimplemented boundary condition. Choose dirichlet_dirichlet or dirichlet_neumann')

    def __repr__(self):
        return f'Grid({self.ng}, {self.length}, {self.Te})'

    def __len__(self):
        return int(self.ng)

    def copy(self):
        return Grid(self.ng, self.length, self.Te)

    # the matrices are kept for API compatibility (attribute A); the solves use the PCR kernel
    def _fill_laplacian_dirichlet(self):
        """pygcpic.py:939-956."""
        ng = self.ng
        self.A = np.zeros((ng, ng))
        for i in range(1, ng - 1):
            self.A[i, i - 1] = 1.0
            self.A[i, i] = -2.0
            self.A[i, i + 1] = 1.0
        self.A[0, 0] = 1.
        self.A[-1, -1] = 1.

    def _fill_laplacian_dirichlet_neumann(self):
        """pygcpic.py:958-978."""
        self._fill_laplacian_dirichlet()
        self.A[-1, -1] = 3.
        self.A[-1, -2] = -4.
        self.A[-1, -3] = 1.

    def weight_particles_to_grid_boltzmann(self, particles, dt):
        """pygcpic.py:841-905.  `particles` is a list of Particle objects (staged as
        structure-of-arrays in one upload) or a ParticleStore already on the device."""
        from pypic_b200 import _lib
        dev = _D.require_cuda()
        if isinstance(particles, ParticleStore):
            st = particles
            x, cs, p2c, act = st.r[0], st.charge_state, st.p2c, st.active
            N = st.N
        else:
            N = len(particles)
            x = _D.to_dev(np.array([p.r[0] for p in particles], dtype=np.float64), dev)
            cs = _D.to_dev(np.array([p.charge_state for p in particles], dtype=np.float64), dev)
            p2c = _D.to_dev(np.array([p.p2c for p in particles], dtype=np.float64), dev)
            act = _D.to_dev(np.array([p.active for p in particles], dtype=np.int8), dev, torch.int8)
        rho = _D.f64(self.ng, dev, True); n = _D.f64(self.ng, dev, True)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        s = _D.stream()
        _lib.call("pic_dev_gc_weight", _D.ptr(x), _D.ptr(cs), _D.ptr(p2c), _D.ptr(act), _D.ptr(rho), _D.ptr(n), N,
                  int(self.ng), float(self.dx), _D.ptr(err), s)
        _D.check_range(err, "Grid.weight_particles_to_grid_boltzmann")
        state = _D.to_dev(np.array([0. if self.n0 is None else self.n0, getattr(self, "p_old", 0.0),
                                    0. if self.n0 is None else 1.]), dev)
        # the staging tensors must outlive the (asynchronous) launch: a temporary passed straight to
        # ptr() is freed at once and its block handed to the next allocation
        d_phi, d_dom = _D.to_dev(self.phi, dev), _D.to_dev(self.domain, dev)
        _lib.call("pic_dev_gc_n0_update", _D.ptr(d_phi), _D.ptr(n), _D.ptr(d_dom),
                  int(self.ng), float(self.Te), float(self.ve), float(self.added_particles), float(dt), _D.ptr(state), s)
        self.rho = rho.cpu().numpy(); self.n = n.cpu().numpy()
        st3 = state.cpu().numpy()
        self.n0 = float(st3[0]); self.p_old = float(st3[1]); self.rho0 = self.n0 * e

    def differentiate_phi_to_E_dirichlet(self):
        """pygcpic.py:907-937."""
        self.E = ops.differentiate(self.phi, self.dx, 3)

    def solve_for_phi_dirichlet(self):
        """pygcpic.py:987-1003."""
        self.phi = ops.poisson_dirichlet(self.rho, self.dx)

    def solve_for_phi_dirichlet_boltzmann(self):
        """pygcpic.py:1005-1053 (exact tridiagonal Newton step instead of bicgstab: same fixed point)."""
        self.phi, self.newton_iterations = ops.newton_boltzmann(self.rho, None, self.dx, self.n0, self.Te, 0, 1e-9, 1000)

    def smooth_rho(self):
        """pygcpic.py:1055-1060."""
        self.rho = ops.smooth(self.rho, 1)

    def solve_for_phi_dirichlet_neumann_boltzmann(self):
        """pygcpic.py:1062-1109."""
        self.phi, self.newton_iterations = ops.newton_boltzmann(self.n, self.phi, self.dx, self.n0, self.Te, 1, 1e-3, 100)

    def reset_added_particles(self):
        self.added_particles = 0

    def add_particles(self, particles):
        """pygcpic.py:1115-1117."""
        self.added_particles += 2 * particles


def _ionizing_decisions(grid, store, dt, ionize_Te, inactive_entry, contrib_entry, Z, source_N, source):
    """One pass of pygcpic.py:1509-1549 after the push for a store that may hold neutrals:
    device probabilities + mid-domain exits, then the sequential event loop on the host (see
    pypic_b200/ionization.py).  Returns (decision tensor, idx_react, r_new, n_del, tallies)."""
    from pypic_b200 import ionization as ION
    N = store.N
    prob, elig, mx, contrib = store.post_push(grid, dt, ION.rates(ionize_Te), Z)
    ev_ion = torch.nonzero(elig[:N] == 1).flatten()
    ev_in = torch.nonzero(inactive_entry[:N] == 1).flatten()
    ev = torch.cat([ev_ion, ev_in])
    kind = torch.cat([torch.zeros_like(ev_ion), torch.ones_like(ev_in)])
    order = torch.argsort(ev)
    ev, kind = ev[order], kind[order]
    ca = contrib[:N].to(torch.int64); ce = contrib_entry[:N].to(torch.int64)
    A = torch.cumsum(ca, 0) - ca                                  # slots before the event, updated state
    Bs = torch.flip(torch.cumsum(torch.flip(ce, [0]), 0), [0])    # slots from the event on, entry state
    h = lambda t: t[ev].cpu().numpy()
    r_new = []
    ionised, new_cs, added, react, deleted = ION.run_events(
        ev.cpu().numpy(), kind.cpu().numpy(), h(prob), h(store.charge_state), h(store.Z), h(store.p2c), h(mx), h(A), h(Bs),
        int(Z), int(source_N), lambda i: r_new.append(next(source)))
    if ionised:
        ii = torch.as_tensor(np.asarray(ionised, dtype=np.int64), device=store.dev)
        store.charge_state[ii] = torch.as_tensor(np.asarray(new_cs), device=store.dev)
        store.invalidate_uniform()
        for a in added:
            grid.add_particles(a)
    dec = torch.zeros(max(N, 1), dtype=torch.int8, device=store.dev)
    if react:
        dec[torch.as_tensor(np.asarray(react, dtype=np.int64), device=store.dev)] = 1
    if deleted:
        dec[torch.as_tensor(np.asarray(deleted, dtype=np.int64), device=store.dev)] = 2
    Zi = h(store.Z)[np.searchsorted(ev.cpu().numpy(), np.asarray(ionised, dtype=np.int64))] if ionised else np.zeros(0)
    tall = dict(ionised_h=int((Zi == 1).sum()), ionised_b=int((Zi == 5).sum()), midexit=int(mx[:N].sum().item()))
    return dec, np.asarray(react, dtype=np.int64), np.array(r_new).reshape(-1, 7), len(deleted), tall


def _global_decisions(grid, store, dt, ionize_Te, inactive_entry, contrib_entry, Z, source_N, source):
    """The decisions of one pass of pygcpic.py:1509-1549 with the particle list SHARDED over the ranks of
    grid.comm (contiguous index ranges in rank order, as every sharded path of this package): every rank
    gathers every rank's event slots (slots inactive at loop entry; ionisation attempts when ionize_Te is
    given), runs the ONE sequential pass of the reference over the global list -- consuming np.random and the
    source generator exactly as a single process would, so all ranks must hold the same stream -- and keeps
    its own share.  Also serves a single rank (event-loop formulation of ParticleStore.decide).
    Returns (decision tensor, own re-activated slots, their 7-vectors, global number deleted, global number
    re-activated, global tallies)."""
    from pypic_b200 import ionization as ION
    comm = grid.comm
    world = comm.world if (comm is not None and comm.enabled) else 1
    rank = comm.rank if world > 1 else 0
    N, dev = store.N, store.dev
    if ionize_Te is not None:
        prob, elig, mx, contrib = store.post_push(grid, dt, ION.rates(ionize_Te), Z)
        ev_ion = torch.nonzero(elig[:N] == 1).flatten()
    else:
        contrib = store.source_ion_flags(Z)
        mx = torch.zeros(max(N, 1), dtype=torch.int8, device=dev)
        prob = torch.zeros(max(N, 1), dtype=torch.float64, device=dev)
        ev_ion = torch.zeros(0, dtype=torch.int64, device=dev)
    ev_in = torch.nonzero(inactive_entry[:N] == 1).flatten()
    ev = torch.cat([ev_ion, ev_in])
    kind = torch.cat([torch.zeros_like(ev_ion), torch.ones_like(ev_in)])
    order = torch.argsort(ev)
    ev, kind = ev[order], kind[order]
    ca = contrib[:N].to(torch.int64); ce = contrib_entry[:N].to(torch.int64)
    A = torch.cumsum(ca, 0) - ca                                  # slots before the event, updated state
    Bs = torch.flip(torch.cumsum(torch.flip(ce, [0]), 0), [0])    # slots from the event on, entry state
    h = lambda t: t[ev].cpu().numpy()
    mine = dict(N=N, ev=ev.cpu().numpy(), kind=kind.cpu().numpy(), prob=h(prob), cs=h(store.charge_state), Z=h(store.Z),
                p2c=h(store.p2c), mx=h(mx), A=h(A), Bs=h(Bs), ca=int(ca.sum().item()) if N else 0,
                ce=int(ce.sum().item()) if N else 0, mxn=int(mx[:N].sum().item()) if N else 0)
    parts = [mine]
    if world > 1:
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=comm.group)
    lo = [0]; Aoff = [0]
    for q in parts:
        lo.append(lo[-1] + q["N"]); Aoff.append(Aoff[-1] + q["ca"])
    Boff = [sum(q["ce"] for q in parts[r + 1:]) for r in range(world)]
    cat = lambda key, add=None: np.concatenate([np.asarray(q[key]) + (0 if add is None else add[r]) for r, q in enumerate(parts)])
    ev_g = cat("ev", lo)
    r_all = []
    ionised, new_cs, added, react, deleted = ION.run_events(
        ev_g, cat("kind"), cat("prob"), cat("cs"), cat("Z"), cat("p2c"), cat("mx"), cat("A", Aoff), cat("Bs", Boff),
        int(Z), int(source_N), lambda i: r_all.append(next(source)))
    a, b = lo[rank], lo[rank + 1]
    own = lambda lst: np.asarray([i - a for i in lst if a <= i < b], dtype=np.int64)
    if ionised:
        sel = [j for j, i in enumerate(ionised) if a <= i < b]
        if sel:
            ii = torch.as_tensor(own(ionised), device=dev)
            store.charge_state[ii] = torch.as_tensor(np.asarray([new_cs[j] for j in sel]), device=dev)
            store.invalidate_uniform()
        for p_ in added:                       # the Boltzmann reference density counts every rank's additions
            grid.add_particles(p_)
    dec = torch.zeros(max(N, 1), dtype=torch.int8, device=dev)
    idx = own(react)
    if len(idx):
        dec[torch.as_tensor(idx, device=dev)] = 1
    dl = own(deleted)
    if len(dl):
        dec[torch.as_tensor(dl, device=dev)] = 2
    r_new = np.array([r_all[j] for j, i in enumerate(react) if a <= i < b]).reshape(-1, 7)
    Zi = cat("Z")[np.searchsorted(ev_g, np.asarray(ionised, dtype=np.int64))] if ionised else np.zeros(0)
    tall = dict(ionised_h=int((Zi == 1).sum()), ionised_b=int((Zi == 5).sum()), midexit=sum(q["mxn"] for q in parts))
    return dec, idx, r_new, len(deleted), len(react), tall


def run_sheath(grid, store, dt, steps, source_N, source, p2c, m, charge_state=1, Z=1, time=0.0, on_step=None,
               ionize_Te=None, event_loop=False):
    """Device-resident time loop with the structure of pic_bca_aps' particle phase
    (pygcpic.py:1486-1563, without the BCA coupling; Monte-Carlo ionisation :350-458 and the
    mid-domain exit of wall-born particles :1530-1541 when ionize_Te is given):

        apply_BCs -> deposit (+Boltzmann n0) -> smooth -> Newton-Boltzmann solve -> E ->
        [gather + Boris + BC fused] -> reactivate-or-delete decision (prefix scan) ->
        re-activation from `source` (host generator, legacy RNG order) -> stable compaction.

    grid: GridDev, store: ParticleStore.  Returns a dict of per-step tallies.

    Sharded runs (GridDev(comm=) over more than one rank, every rank holding a contiguous index range of the
    reference's particle list, rank order = index order, and the SAME np.random / source-generator state): the
    deposits are all-reduced by GridDev, the order-dependent decisions are taken over the global event list on
    every rank (_global_decisions), and the per-step tallies are global.  event_loop=True forces that
    formulation on a single rank (tests)."""
    out = dict(length=[], hits=[], deleted=[], reactivated=[], n0=[], ekin=[], angle=[], ionised_h=[], ionised_b=[],
               midexit=[])
    comm = getattr(grid, "comm", None)
    sharded = comm is not None and comm.enabled and comm.world > 1
    for _ in range(int(steps)):
        time += dt
        if grid.have_fused_n:
            # a species-uniform store: the previous step's push kernel already deposited the
            # survivors at their new positions (and flagged everything outside the walls), the
            # re-activated slots were added by ParticleStore.reactivate
            grid.finish_fused_deposit(charge_state, dt)
        else:
            store.apply_BCs_dirichlet(grid)
            grid.weight_particles_to_grid_boltzmann(store, dt)
        grid.smooth_rho()
        grid.reset_added_particles()
        grid.solve_for_phi_dirichlet_boltzmann()
        grid.differentiate_phi_to_E_dirichlet()
        inactive_entry = (store.active != 1).to(torch.int8)
        contrib_entry = store.source_ion_flags(Z)
        # The fused kernel may deposit the survivors for the NEXT step only if nothing changes them in
        # between: with Monte-Carlo ionisation the charge states change, wall-born particles leave
        # mid-domain (pygcpic.py:1530-1541) and the deposit needs every particle's own charge state
        # (:871-883), so that case deposits in its own pass at the top of the next step.
        hits = store.push_6D(dt, grid, deposit=ionize_Te is None)
        ke, ang, _ = store.wall_hit_tallies()
        if sharded or event_loop:
            dec, idx, r_new, n_del, n_react, tall = _global_decisions(grid, store, dt, ionize_Te, inactive_entry, contrib_entry,
                                                                      Z, source_N, source)
            if len(idx):
                store.reactivate(idx, r_new, p2c, m, charge_state, Z, time, grid)
            for _k in range(n_react - len(idx)):       # the other ranks' re-activations count in the reference density
                grid.add_particles(p2c)
            if ionize_Te is not None:
                for k_, v_ in tall.items():
                    out[k_].append(v_)
            if sharded:
                import torch.distributed as dist
                got = [None] * comm.world
                dist.all_gather_object(got, (hits, ke, ang, store.N - int((dec[:store.N] == 2).sum().item())), group=comm.group)
                hits = sum(g_[0] for g_ in got)
                ke = np.concatenate([np.asarray(g_[1]) for g_ in got]); ang = np.concatenate([np.asarray(g_[2]) for g_ in got])
                n_global_after = sum(g_[3] for g_ in got)
        elif ionize_Te is None:
            contrib_after = store.source_ion_flags(Z)
            dec, n_react, n_del = store.decide(inactive_entry, contrib_entry, contrib_after, source_N)
            if n_react:
                idx = torch.nonzero(dec[:store.N] == 1).flatten().cpu().numpy()
                r_new = np.array([next(source) for _ in idx])
                store.reactivate(idx, r_new, p2c, m, charge_state, Z, time, grid)
        else:
            # Monte-Carlo ionisation + mid-domain exits (ionize_Te = electron temperature [K] of
            # attempt_*_ionization's np.interp): decisions by the host event loop, index order
            dec, idx, r_new, n_del, tall = _ionizing_decisions(grid, store, dt, ionize_Te, inactive_entry, contrib_entry,
                                                               Z, source_N, source)
            n_react = len(idx)
            if n_react:
                store.reactivate(idx, r_new, p2c, m, charge_state, Z, time, grid)
            for k_, v_ in tall.items():
                out[k_].append(v_)
        store.compact(dec)
        out["length"].append(n_global_after if sharded else store.N); out["hits"].append(hits); out["deleted"].append(n_del)
        out["reactivated"].append(n_react); out["n0"].append(grid.n0); out["ekin"].append(ke); out["angle"].append(ang)
        if on_step is not None:
            n_before = store.N
            on_step(grid, store)
            if store.N != n_before or store._uniform is False:
                grid.have_fused_n = False          # the callback changed the store: deposit it afresh next step
    store.check(); grid.check()
    return out
