"""Device-resident structure-of-arrays mirror of pygcpic.py's Particle list and Grid.

ParticleStore holds what the reference keeps per Particle object (pygcpic.py:77-112):
the 7-vector r = [x,y,z,vx,vy,vz,t] as seven fp64 arrays, charge_state, m, p2c (fp64),
Z (int32) and the int8 flags active / at_wall / from_wall.  GridDev holds the Grid arrays
(pygcpic.py:781-807) and the Boltzmann reference-density state.  Every method launches
CUDA kernels through the C ABI; nothing is computed on the host.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, device as D

epsilon0 = 8.854e-12
e = 1.602e-19
mp = 1.67e-27
me = 9.11e-31
kb = 1.38e-23


class GridDev:
    def __init__(self, ng, length, Te, bc="dirichlet-dirichlet", device=None, comm=None):
        """comm (pypic_b200.dist.Comm): particle decomposition over ranks -- every rank deposits its
        own particles and the grid moments [rho | n] are all-reduced; the field
        solve is replicated."""
        assert ng > 1, "Number of grid points must be greater than 1"
        assert length > 0.0, "Length must be greater than 0"
        if not isinstance(bc, str):
            raise TypeError("bc must be a string")
        if bc not in ("dirichlet-dirichlet", "dirichlet-neumann"):
            raise ValueError("Unimplemented boundary condition. Choose dirichlet_dirichlet or dirichlet_neumann")
        self.dev = D.require_cuda(device)
        self.comm = comm
        self.ng, self.length, self.Te, self.bc = int(ng), float(length), float(Te), bc
        self.domain_h = np.linspace(0.0, length, ng)
        self.dx = float(self.domain_h[1] - self.domain_h[0])
        self.ve = float(np.sqrt(8. / np.pi * kb * self.Te / me))
        dev = self.dev
        self.domain = D.to_dev(self.domain_h, dev)
        self.rho = D.f64(ng, dev, True); self.phi = D.f64(ng, dev, True)
        self.E = D.f64(ng, dev, True); self.n = D.f64(ng, dev, True)
        self.state = D.f64(3, dev, True)            # n0, p_old, initialised
        self.work = D.f64(13 * ng, dev, True)
        self.iters = torch.zeros(1, dtype=torch.int32, device=dev)
        self.range_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.added_particles = 0.0
        self.newton_iterations = 0
        self.n_acc = D.f64(ng, dev, True)           # number density deposited by the fused push
        self.rho_acc = D.f64(ng, dev, True)         # charge density, the same (mixed-species stores only)
        self.have_fused_n = False
        self._fused_mixed = False

    # -- reference-density state ------------------------------------------------------
    @property
    def n0(self):
        s = D.read_f64(self.state, 3)
        return None if s[2] == 0.0 else float(s[0])

    @n0.setter
    def n0(self, v):
        s = D.read_f64(self.state, 3)
        if v is None:
            s[:] = 0.0
        else:
            s[0] = v; s[2] = 1.0
        self.state.copy_(torch.as_tensor(s))

    @property
    def rho0(self):
        v = self.n0
        return None if v is None else v * e

    def reset_added_particles(self):
        self.added_particles = 0.0

    def add_particles(self, particles):
        self.added_particles += 2. * particles      # pygcpic.py:1115-1117

    # -- deposit -----------------------------------------------------------------------
    def weight_particles_to_grid_boltzmann(self, store, dt):
        """pygcpic.py:841-905."""
        st = D.stream()
        self.rho.zero_(); self.n.zero_()
        _lib.call("pic_dev_gc_weight", D.ptr(store.r[0]), D.ptr(store.charge_state), D.ptr(store.p2c),
                  D.ptr(store.active), D.ptr(self.rho), D.ptr(self.n), store.N, self.ng, self.dx,
                  D.ptr(self.range_err), st)
        if self.comm is not None:
            self.comm.allreduce_sum(self.rho); self.comm.allreduce_sum(self.n)
        _lib.call("pic_dev_gc_n0_update", D.ptr(self.phi), D.ptr(self.n), D.ptr(self.domain), self.ng, self.Te, self.ve,
                  float(self.added_particles), float(dt), D.ptr(self.state), st)

    # -- fused deposit ----------------------------------------------------------------------
    def begin_fused_deposit(self, mixed=False):
        """mixed: the push kernel deposits rho per particle as well (a store with several species);
        otherwise rho = charge_state*e*n at finish_fused_deposit."""
        self.n_acc.zero_()
        if mixed:
            self.rho_acc.zero_()
        self.have_fused_n = True
        self._fused_mixed = bool(mixed)

    def promote_fused_to_mixed(self, charge_state):
        """A pending species-uniform deposit (rho implied by n) becomes a two-accumulator one: a particle
        of another species is about to be added to it."""
        if self.have_fused_n and not self._fused_mixed:
            _lib.call("pic_dev_gc_uniform_finish", D.ptr(self.n_acc), D.ptr(self.n), D.ptr(self.rho_acc), self.ng,
                      float(charge_state), D.stream())
            self._fused_mixed = True

    def deposit_slots(self, store, idx, p2c, charge_state=None):
        """Adds the density of the slots idx (int64 device tensor) -- the particles
        re-activated after the fused push -- to the pending deposit."""
        if idx.numel():
            _lib.call("pic_dev_gc_deposit_idx", D.ptr(store.r[0]), D.ptr(idx), idx.numel(), float(p2c), self.dx, self.ng,
                      D.ptr(self.n_acc), D.ptr(self.range_err), D.stream())
            if self._fused_mixed:
                if charge_state is None:
                    raise _lib.PicError(_lib.PIC_ERR_ARG, "deposit_slots: a mixed-species pending deposit needs the charge state")
                # charge_state*e*p2c in the reference's product order (pygcpic.py:871-873)
                _lib.call("pic_dev_gc_deposit_idx", D.ptr(store.r[0]), D.ptr(idx), idx.numel(),
                          float(charge_state) * 1.602e-19 * float(p2c), self.dx, self.ng, D.ptr(self.rho_acc),
                          D.ptr(self.range_err), D.stream())

    def finish_fused_deposit(self, charge_state, dt):
        """n, rho from the density the fused push deposited (+ re-activated slots), then the
        Boltzmann reference-density update of pygcpic.py:889-904."""
        st = D.stream()
        if self.comm is not None:
            self.comm.allreduce_sum(self.n_acc)
            if self._fused_mixed:
                self.comm.allreduce_sum(self.rho_acc)
        if self._fused_mixed:
            self.n.copy_(self.n_acc); self.rho.copy_(self.rho_acc)
        else:
            _lib.call("pic_dev_gc_uniform_finish", D.ptr(self.n_acc), D.ptr(self.n), D.ptr(self.rho), self.ng,
                      float(charge_state), st)
        _lib.call("pic_dev_gc_n0_update", D.ptr(self.phi), D.ptr(self.n), D.ptr(self.domain), self.ng, self.Te, self.ve,
                  float(self.added_particles), float(dt), D.ptr(self.state), st)
        self.have_fused_n = False

    def smooth_rho(self):
        out = torch.empty_like(self.rho)
        _lib.call("pic_dev_smooth", D.ptr(self.rho), D.ptr(out), self.ng, 1, D.stream())
        self.rho = out

    # -- field solves --------------------------------------------------------------------
    def solve_for_phi_dirichlet(self):
        _lib.call("pic_dev_poisson_dirichlet", D.ptr(self.rho), D.ptr(self.phi), self.ng, self.dx, D.ptr(self.work),
                  D.stream())

    def solve_for_phi_dirichlet_boltzmann(self):
        n0 = self.n0
        _lib.call("pic_dev_newton_boltzmann", D.ptr(self.rho), D.ptr(self.phi), self.ng, self.dx, float(n0), self.Te, 0,
                  1e-9, 1000, D.ptr(self.iters), D.stream())

    def solve_for_phi_dirichlet_neumann_boltzmann(self):
        n0 = self.n0
        _lib.call("pic_dev_newton_boltzmann", D.ptr(self.n), D.ptr(self.phi), self.ng, self.dx, float(n0), self.Te, 1,
                  1e-3, 100, D.ptr(self.iters), D.stream())

    def solve_for_phi(self):
        if self.bc == "dirichlet-dirichlet":
            self.solve_for_phi_dirichlet_boltzmann()
        else:
            self.solve_for_phi_dirichlet_neumann_boltzmann()

    def differentiate_phi_to_E_dirichlet(self):
        _lib.call("pic_dev_differentiate", D.ptr(self.phi), D.ptr(self.E), self.ng, self.dx, 3, D.stream())

    def check(self):
        D.check_range(self.range_err, "pygcpic grid deposit")


class ParticleStore:
    FIELDS_F64 = ("charge_state", "m", "p2c")
    FLAGS = ("active", "at_wall", "from_wall")

    def __init__(self, N, B=(0., 0., 0.), Eyz=(0., 0.), device=None):
        self.dev = D.require_cuda(device)
        self.N = int(N)
        n = max(self.N, 1)
        dev = self.dev
        self.r = [D.f64(n, dev, True) for _ in range(7)]
        self.charge_state = D.f64(n, dev, True); self.m = D.f64(n, dev, True); self.p2c = D.f64(n, dev, True)
        self.Z = torch.zeros(n, dtype=torch.int32, device=dev)
        self.active = torch.ones(n, dtype=torch.int8, device=dev)
        self.at_wall = torch.zeros(n, dtype=torch.int8, device=dev)
        self.from_wall = torch.zeros(n, dtype=torch.int8, device=dev)
        self.hit_flag = torch.zeros(n, dtype=torch.int8, device=dev)
        self.B = tuple(float(b) for b in B)
        self.Eyz = tuple(float(v) for v in Eyz)
        self.hit_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.range_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.mode = 0
        self._uniform = False                       # False: unknown, None: not uniform, tuple: (cs, m, p2c)
        # carry_yzt = False ("lean" store): the fused Boris kernel streams x, vx, vy, vz only (64 B per
        # particle-step instead of 112 B): y and z are not advanced (nothing on the path reads them: the
        # field depends on x alone) and the per-particle clock r[6] is kept implicitly -- it equals the
        # time of the last push for every active particle and is written at the moment a particle dies
        self.carry_yzt = True
        self.time = 0.0                             # time of the last push_6D (lean stores: the active particles' clock)

    # -- construction / export -----------------------------------------------------------
    @classmethod
    def from_arrays(cls, r, charge_state, m, p2c, Z=None, active=None, at_wall=None, from_wall=None, B=(0, 0, 0),
                    Eyz=(0, 0), device=None):
        r = np.asarray(r, dtype=np.float64)
        s = cls(r.shape[0], B, Eyz, device)
        N = s.N                     # the arrays hold max(N, 1) slots so that an empty store still has pointers
        for c in range(7):
            s.r[c][:N].copy_(torch.as_tensor(np.ascontiguousarray(r[:, c])))
        for name, val in (("charge_state", charge_state), ("m", m), ("p2c", p2c)):
            getattr(s, name)[:N].copy_(torch.as_tensor(np.broadcast_to(np.asarray(val, dtype=np.float64), (N,)).copy()))
        if Z is not None:
            s.Z[:N].copy_(torch.as_tensor(np.broadcast_to(np.asarray(Z, dtype=np.int32), (N,)).copy()))
        for name, val in (("active", active), ("at_wall", at_wall), ("from_wall", from_wall)):
            if val is not None:
                getattr(s, name)[:N].copy_(torch.as_tensor(np.asarray(val).astype(np.int8)))
        return s

    FUSED_MIN = 16384                               # one chunk of the v2 kernel
    MIXED_MAX_NG = 6900                             # field tile + two deposit windows + ring of the mixed kernel in 227 KB

    def uniform(self):
        """(charge_state, m, p2c) if every slot holds the same values (then the push can take
        them as scalars instead of streaming 24 B/particle), else None.  Cached; methods that
        write these arrays invalidate or re-check it."""
        if self._uniform is False:
            n = self.N
            if n == 0:
                self._uniform = None
            else:
                vals = []
                for a in (self.charge_state, self.m, self.p2c):
                    lo, hi = torch.aminmax(a[:n])
                    vals.append((float(lo), float(hi)))
                self._uniform = tuple(v[0] for v in vals) if all(v[0] == v[1] for v in vals) else None
        return self._uniform

    def invalidate_uniform(self):
        """Call after writing charge_state / m / p2c tensors directly."""
        self._uniform = False

    def r_host(self):
        r = np.stack([c[:self.N].cpu().numpy() for c in self.r], 1)
        if not self.carry_yzt:
            act = self.active[:self.N].cpu().numpy() == 1
            r[:, 1:3] = np.nan                      # not tracked by a lean store
            r[act, 6] = self.time                   # the clock of the active particles is the time of the last push
        return r

    def flags_host(self):
        return {k: getattr(self, k)[:self.N].cpu().numpy() for k in self.FLAGS}

    def _params(self, grid, dt=0.0, flags=0):
        return _lib.GCParams(self.N, grid.ng if grid is not None else 2, int(flags), grid.dx if grid is not None else 1.0,
                             float(dt), grid.length if grid is not None else 1.0, (C.c_double * 3)(*self.B),
                             (C.c_double * 2)(*self.Eyz))

    def _r7(self):
        return _lib.R7(*[D.ptr(c) for c in self.r])

    # -- per-step operations -------------------------------------------------------------
    def apply_BCs_dirichlet(self, grid):
        """pygcpic.py:668-689 for every particle."""
        _lib.call("pic_dev_gc_apply_bcs", D.ptr(self.r[0]), D.ptr(self.active), D.ptr(self.at_wall), self.N, grid.length,
                  D.stream())

    def gather(self, grid):
        """interpolate_electric_field_dirichlet (mirrored weights) -> numpy E_x per particle."""
        out = D.f64(max(self.N, 1), self.dev)
        _lib.call("pic_dev_gc_interpolate", D.ptr(grid.E), D.ptr(self.r[0]), D.ptr(out), self.N, grid.ng, grid.dx,
                  D.ptr(self.range_err), D.stream())
        return out[:self.N].cpu().numpy()

    def push_6D(self, dt, grid, deposit=False, time=None):
        """Fused interpolate_electric_field_dirichlet + push_6D + apply_BCs_dirichlet for all
        active particles (pygcpic.py:1500-1502).  Returns the number of wall hits.

        A species-uniform store takes the TMA-ring kernel; with deposit=True that kernel also
        deposits the number density of the survivors at their new positions into grid.n_acc
        (the next step's weight_particles_to_grid_boltzmann, see GridDev.finish_fused_deposit).
        time: the simulation time after this push (default: the store's own clock + dt)."""
        P = self._params(grid, dt)
        r7 = self._r7()
        self.hit_count.zero_()
        self.time = float(time) if time is not None else self.time + float(dt)
        fused = self.N >= self.FUSED_MIN and grid.ng >= 8
        u = self.uniform() if fused else None
        if not fused and not self.carry_yzt:
            raise _lib.PicError(_lib.PIC_ERR_ARG, "a lean store (carry_yzt=False) is served by the fused kernels only "
                                "(N >= %d)" % self.FUSED_MIN)
        if u is not None:
            if deposit:
                grid.begin_fused_deposit()
            _lib.call("pic_dev_gc_push_boris_uniform2", C.byref(P), C.byref(r7), u[0], u[1], u[2],
                      0 if self.carry_yzt else 1, self.time, D.ptr(self.active), D.ptr(self.at_wall), D.ptr(self.hit_flag),
                      D.ptr(grid.E), D.ptr(grid.n_acc) if deposit else None, D.ptr(self.hit_count), D.ptr(self.range_err),
                      D.stream())
            return int(D.read_raw(self.hit_count, 1, np.int64)[0])
        if fused and (grid.ng <= self.MIXED_MAX_NG or not self.carry_yzt):
            # several species in one list: per-particle charge_state, m, p2c ride the ring; n AND rho deposited
            if deposit:
                grid.begin_fused_deposit(mixed=True)
            _lib.call("pic_dev_gc_push_boris_mixed", C.byref(P), C.byref(r7), D.ptr(self.charge_state), D.ptr(self.m),
                      D.ptr(self.p2c), 0 if self.carry_yzt else 1, self.time, D.ptr(self.active), D.ptr(self.at_wall),
                      D.ptr(self.hit_flag), D.ptr(grid.E), D.ptr(grid.n_acc) if deposit else None,
                      D.ptr(grid.rho_acc) if deposit else None, D.ptr(self.hit_count), D.ptr(self.range_err), D.stream())
            return int(D.read_raw(self.hit_count, 1, np.int64)[0])
        _lib.call("pic_dev_gc_push_boris", C.byref(P), C.byref(r7), D.ptr(self.charge_state), D.ptr(self.m),
                  D.ptr(self.active), D.ptr(self.at_wall), D.ptr(self.hit_flag), D.ptr(grid.E), D.ptr(self.hit_count),
                  D.ptr(self.range_err), D.stream())
        return int(D.read_raw(self.hit_count, 1, np.int64)[0])

    def push_6D_given_E(self, dt, Ex):
        """Particle.push_6D alone (pygcpic.py:460-507): E_x per particle already gathered (device
        tensor), no boundary test."""
        P = self._params(None, dt, flags=3)
        r7 = self._r7()
        _lib.call("pic_dev_gc_push_boris", C.byref(P), C.byref(r7), D.ptr(self.charge_state), D.ptr(self.m),
                  D.ptr(self.active), D.ptr(self.at_wall), None, D.ptr(Ex), None, D.ptr(self.range_err), D.stream())

    def push_GC_given_E(self, dt, Ex):
        """Particle.push_GC (pygcpic.py:598-645) with E_x per particle already gathered."""
        P = self._params(None, dt, flags=1)
        r7 = self._r7()
        _lib.call("pic_dev_gc_push_rk4", C.byref(P), C.byref(r7), D.ptr(self.charge_state), D.ptr(self.m),
                  D.ptr(self.active), D.ptr(Ex), D.ptr(self.range_err), D.stream())

    def transform_6D_to_GC(self):
        P = self._params(None)
        r7 = self._r7()
        _lib.call("pic_dev_gc_to_gc", C.byref(P), C.byref(r7), D.ptr(self.charge_state), D.ptr(self.m),
                  D.ptr(self.active), D.stream())
        self.mode = 1

    def transform_GC_to_6D(self, a):
        """a: (N,3) uniform(0,1) draws (pygcpic.py:583), made on the host in particle order."""
        a = np.asarray(a, dtype=np.float64)
        P = self._params(None)
        r7 = self._r7()
        da = [D.to_dev(a[:, c], self.dev) for c in range(3)]
        _lib.call("pic_dev_gc_to_6d", C.byref(P), C.byref(r7), D.ptr(self.charge_state), D.ptr(self.m),
                  D.ptr(self.active), D.ptr(da[0]), D.ptr(da[1]), D.ptr(da[2]), D.stream())
        torch.cuda.current_stream().synchronize()
        self.mode = 0

    RK4_UNIFORM = True                              # class switch for tests (v1 kernel when False)

    def push_GC(self, dt, grid=None):
        P = self._params(grid, dt)
        r7 = self._r7()
        u = self.uniform() if self.RK4_UNIFORM else None
        if u is not None and u[0] != 0.0 and any(self.B):
            _lib.call("pic_dev_gc_push_rk4_uniform", C.byref(P), C.byref(r7), u[0], u[1], D.ptr(self.active),
                      D.ptr(grid.E) if grid is not None else None, D.ptr(self.range_err), D.stream())
            return
        _lib.call("pic_dev_gc_push_rk4", C.byref(P), C.byref(r7), D.ptr(self.charge_state), D.ptr(self.m),
                  D.ptr(self.active), D.ptr(grid.E) if grid is not None else None, D.ptr(self.range_err), D.stream())

    def post_push(self, grid, dt, rate4, source_Z):
        """pic_dev_gc_post_push: ionisation eligibility/probability, mid-domain exit of wall-born
        particles (applied to `active`), deterministic source-ion contribution.  Returns device
        tensors (prob f64, eligible i8, midexit i8, contrib i8)."""
        n = max(self.N, 1)
        prob = D.f64(n, self.dev)
        elig = torch.empty(n, dtype=torch.int8, device=self.dev)
        mx = torch.empty_like(elig); contrib = torch.empty_like(elig)
        _lib.call("pic_dev_gc_post_push", D.ptr(self.r[0]), D.ptr(self.p2c), D.ptr(self.charge_state), D.ptr(self.Z),
                  D.ptr(self.from_wall), D.ptr(self.active), D.ptr(grid.n), grid.ng, grid.dx, float(dt), grid.length,
                  C.byref((C.c_double * 4)(*rate4)), int(source_Z), D.ptr(prob), D.ptr(elig), D.ptr(mx), D.ptr(contrib),
                  self.N, D.ptr(self.range_err), D.stream())
        return prob, elig, mx, contrib

    # -- reactivate-or-delete + compaction -------------------------------------------------
    def source_ion_flags(self, source_Z):
        """int8 flag: Z==source and active and charge_state>0 (pygcpic.py:1544)."""
        return ((self.Z == int(source_Z)) & (self.active == 1) & (self.charge_state > 0)).to(torch.int8)

    def decide(self, inactive_entry, contrib_entry, contrib_after, source_N):
        """Order-dependent rule of pygcpic.py:1543-1549.  Returns (decision int8 tensor,
        n_reactivated, n_deleted)."""
        N = self.N
        dec = torch.zeros(max(N, 1), dtype=torch.int8, device=self.dev)
        idx = torch.empty(max(N, 1), dtype=torch.int32, device=self.dev)
        base = torch.empty(max(N, 1), dtype=torch.int32, device=self.dev)
        scratch = torch.zeros(4 + 2 * (N // 2048 + 2), dtype=torch.int64, device=self.dev)
        _lib.call("pic_dev_gc_decide", D.ptr(inactive_entry), D.ptr(contrib_entry), D.ptr(contrib_after), D.ptr(dec),
                  N, int(source_N), D.ptr(idx), D.ptr(base), D.ptr(scratch), D.stream())
        s = D.read_raw(scratch, 4, np.int64)
        return dec, int(s[2]), int(s[3])

    def reactivate(self, where_idx, r_new, p2c, m, charge_state, Z, time, grid):
        """Particle.reactivate (pygcpic.py:691-720) for the slots in where_idx (host int
        array, index order) with host-drawn 7-vectors r_new (len(where_idx),7)."""
        if len(where_idx) == 0:
            return
        idx = torch.as_tensor(np.asarray(where_idx, dtype=np.int64), device=self.dev)
        r_new = np.asarray(r_new, dtype=np.float64).copy()
        r_new[:, 6] = time
        rn = torch.as_tensor(r_new, device=self.dev)
        for c in range(7):
            self.r[c][idx] = rn[:, c]
        self.p2c[idx] = float(p2c); self.m[idx] = float(m); self.charge_state[idx] = float(charge_state)
        self.Z[idx] = int(Z)
        self.active[idx] = 1; self.at_wall[idx] = 0; self.from_wall[idx] = 0
        self.hit_flag[idx] = 0
        if self._uniform not in (False, None) and self._uniform != (float(charge_state), float(m), float(p2c)):
            grid.promote_fused_to_mixed(self._uniform[0])
            self._uniform = None
        if grid.have_fused_n:                      # the fused push has already deposited the survivors
            grid.deposit_slots(self, idx, p2c, charge_state)
        for _ in range(len(where_idx)):
            grid.add_particles(p2c)

    def compact(self, decision):
        """Stable (order-preserving) removal of the slots flagged 2 (pygcpic.py:1552-1563)."""
        N = self.N
        idx = torch.empty(max(N, 1), dtype=torch.int32, device=self.dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=self.dev)
        bc = torch.zeros(2 * (N // 2048 + 2), dtype=torch.int64, device=self.dev)
        st = D.stream()
        _lib.call("pic_dev_compact_flags", D.ptr(decision), N, 2, D.ptr(idx), D.ptr(cnt), D.ptr(bc), st)
        M = int(D.read_raw(cnt, 1, np.int64)[0])
        if M == N:
            return 0

        def g64(src):
            dst = torch.empty(max(M, 1), dtype=torch.float64, device=self.dev)
            _lib.call("pic_dev_gather_f64", D.ptr(src), D.ptr(idx), D.ptr(dst), M, st)
            return dst

        def g8(src):
            dst = torch.empty(max(M, 1), dtype=torch.int8, device=self.dev)
            _lib.call("pic_dev_gather_i8", D.ptr(src), D.ptr(idx), D.ptr(dst), M, st)
            return dst
        self.r = [g64(c) for c in self.r]
        self.charge_state, self.m, self.p2c = g64(self.charge_state), g64(self.m), g64(self.p2c)
        self.Z = self.Z[idx[:M].long()].contiguous() if M else self.Z[:1]
        self.active, self.at_wall, self.from_wall = g8(self.active), g8(self.at_wall), g8(self.from_wall)
        self.hit_flag = torch.zeros(max(M, 1), dtype=torch.int8, device=self.dev)
        if getattr(self, "perm", None) is not None:
            self.perm = self.perm[idx[:M].long()]
        removed = N - M
        self.N = M
        return removed

    def sort_by_cell(self, grid, track=False):
        """Re-orders the store by grid cell (counting sort of x with the slot index as payload,
        then gathers of every other array): keeps a warp's particles inside the deposit window of
        the fused kernel.  The order inside a cell is unspecified; with track=True self.perm
        composes the permutations so that slot s holds the particle that was at self.perm[s] when
        the store was created."""
        N = self.N
        if N < 2:
            return
        dev = self.dev
        # a store sorted a few steps ago is NEARLY sorted: the global-cursor path of the counting sort (flags
        # bit5) is then the faster one
        P = _lib.DDParams(N, N, grid.ng, 32 if getattr(self, "_sorted_once", False) else 0, grid.dx, 1.0, grid.length, 1.0,
                          (C.c_double * 2)(0., 0.), (C.c_double * 2)(1., 1.))
        self._sorted_once = True
        new = lambda dt=torch.float64: torch.empty(N, dtype=dt, device=dev)
        xs, vxs, vys, vzs = new(), new(), new(), new()
        idx = new(torch.int32)
        counts = torch.zeros(D.sort_counts_size(grid.ng), dtype=torch.int32, device=dev)
        st = D.stream()
        # x and the three velocity components travel THROUGH the scatter of the counting sort (runs of
        # coalesced writes); only the remaining, smaller arrays are gathered through the permutation
        _lib.call("pic_dev_sort_by_cell_payload", C.byref(P), D.ptr(self.r[0]), D.ptr(self.r[3]), D.ptr(self.r[4]),
                  D.ptr(self.r[5]), D.ptr(xs), D.ptr(vxs), D.ptr(vys), D.ptr(vzs), D.ptr(idx), D.ptr(counts), st)
        # a species-uniform store holds the same charge_state / m / p2c in every slot: permuting
        # those three arrays would be the identity (saves 48 of ~180 B/particle of the permutation)
        uni = self.uniform() is not None
        comps = [1, 2, 6] if self.carry_yzt else [6]                       # a lean store does not track y, z
        f_src = [self.r[c] for c in comps] + ([] if uni else [self.charge_state, self.m, self.p2c])
        f_dst = [new() for _ in f_src]
        b_src = [self.active, self.at_wall, self.from_wall, self.hit_flag]
        b_dst = [new(torch.int8) for _ in b_src]
        z_dst = new(torch.int32)
        arr = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        _lib.call("pic_dev_soa_permute", D.ptr(idx), N, arr(f_src), arr(f_dst), len(f_src), arr([self.Z]), arr([z_dst]), 1,
                  arr(b_src), arr(b_dst), len(b_src), st)
        newr = list(self.r)
        newr[0], newr[3], newr[4], newr[5] = xs, vxs, vys, vzs
        for c, t in zip(comps, f_dst):
            newr[c] = t
        self.r = newr
        if not uni:
            self.charge_state, self.m, self.p2c = f_dst[len(comps):]
        self.Z = z_dst
        self.active, self.at_wall, self.from_wall, self.hit_flag = b_dst
        if track:
            prev = getattr(self, "perm", None)
            self.perm = idx.long() if prev is None else prev[idx.long()]

    def append(self, other):
        """particles += new_particles (pygcpic.py:1624)."""
        keep = self.N
        cat = lambda a, b, n1, n2: torch.cat([a[:n1], b[:n2]])
        self.r = [cat(a, b, keep, other.N) for a, b in zip(self.r, other.r)]
        for name in self.FIELDS_F64 + self.FLAGS + ("Z", "hit_flag"):
            setattr(self, name, cat(getattr(self, name), getattr(other, name), keep, other.N))
        self.N = keep + other.N
        self._uniform = False

    def wall_hit_tallies(self):
        """kinetic_energy/e and angle w.r.t. the wall (pygcpic.py:262-275, 228-259) of the
        particles absorbed by the last push_6D, in index order (host arrays)."""
        idx = torch.nonzero(self.hit_flag[:self.N] == 1).flatten()
        if idx.numel() == 0:
            return np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int64)
        v = np.stack([self.r[c][idx].cpu().numpy() for c in (3, 4, 5)], 1)
        m = self.m[idx].cpu().numpy()
        speed = np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2 + v[:, 2] ** 2)
        ke = 0.5 * m * speed ** 2 / e
        ang = np.arctan2(np.sqrt(v[:, 1] ** 2 + v[:, 2] ** 2), np.abs(v[:, 0])) * 180. / np.pi
        return ke, ang, idx.cpu().numpy()

    def check(self):
        D.check_range(self.range_err, "pygcpic particle kernels")
