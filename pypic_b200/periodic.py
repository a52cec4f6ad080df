"""Device-resident drivers of the two periodic codes:

* PeriodicImplicitSim -- pypic.py's implicit Crank-Nicolson / Picard loop
  (pypic.particle_push_p, pypic.py:216-300) for one species;
* ExplicitSim -- PIC_L.py's explicit leapfrog loop with a Poisson solve every step
  (PIC_L.main, PIC_L.py:762-768).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, device as D
from .dist import Comm, local_split, shard_range

epsilon0 = 8.854E-12
e = 1.602E-19
mp = 1.67E-27
me = 9.11E-31
kb = 1.38E-23


class PeriodicImplicitSim:
    def __init__(self, N, Ng, dx, dt, L, p2c, q=-e, m=me, tol=1e-3, maxiter=20, deposit="warp", comm=None,
                 device=None, sort_every=0, track_order=True):
        self.dev = D.require_cuda(device)
        self.comm = comm if comm is not None else Comm()
        self.N_global = int(N)
        self.start, self.stop = shard_range(N, self.comm.rank, self.comm.world)
        self.N = self.stop - self.start
        self.Ng, self.dx, self.dt, self.L = int(Ng), float(dx), float(dt), float(L)
        self.p2c = float(int(p2c))           # numba's int32 signature truncates p2c
        self.p2c_raw = float(p2c)            # the Python-level diagnostics use the untruncated value
        self.tol, self.maxiter = float(tol), int(maxiter)
        # deposit: "window" = TMA-staged private-window kernel (needs a store sorted by cell: set
        # sort_every), "warp" = grid-stride kernel with warp pre-reduced shared-memory atomics (any
        # particle order; the default of the drop-in modules), "atomic" = one atomicAdd per contribution
        if sort_every and deposit == "warp":
            deposit = "window"
        # bit1: the store keeps the UNWRAPPED x1 of the previous step and the kernels apply
        # ``x % L`` (pypic.py:277) when they load it -- saves a 16 B/particle pass per step
        # "window-big" forces the large-grid build of the window kernel (per-warp field windows instead of the
        # whole-grid tile; taken automatically when the grid does not fit shared memory, Ng >~ 9000)
        # "window-det": the REPRODUCIBLE build (flags bit7, as SheathSim's): integer merges of the currents on fixed-point
        # words, stable radix sort (with the original index as its payload when track_order), fixed-order kinetic
        # energy -- two runs give bit-identical output
        flags = {"window": 0, "window-big": 16, "warp": 4, "atomic": 1 | 4, "window-det": 128}[deposit] | 2
        self.det = deposit == "window-det"
        self.sort_every = int(sort_every)
        self.t = 0
        self.perm = None                     # original index of the particle in each slot (after sorting)
        self.track_order = bool(track_order)  # False: do not carry the original index through the sorts (download()
                                              # then returns the particles in store order)
        # q, m: scalars (one species, the reference's only use) or per-particle arrays (the signature of
        # particle_push_p, pypic.py:248): arrays take the grid-stride kernel, unsorted, full iterations
        self.qm_arrays = None
        if (np.ndim(q) or np.ndim(m)) and self.det:
            raise ValueError("deposit='window-det' is built for one species (scalar q, m)")
        if np.ndim(q) or np.ndim(m):
            qa = np.broadcast_to(np.asarray(q, dtype=np.float64), (self.N_global,))[self.start:self.stop].copy()
            ma = np.broadcast_to(np.asarray(m, dtype=np.float64), (self.N_global,))[self.start:self.stop].copy()
            self.qm_arrays = (D.to_dev(qa, self.dev), D.to_dev(ma, self.dev))
            self.sort_every = 0
            flags = (flags | 4) & ~1
            q, m = float(qa[0]) if len(qa) else -e, float(ma[0]) if len(ma) else me
        self.params = _lib.PypicParams(self.N, self.Ng, flags, self.dx, self.dt, self.L, self.p2c, float(q), float(m))
        dev, n, g = self.dev, max(self.N, 1), self.Ng
        self.x0 = D.f64(n, dev, True); self.v0 = D.f64(n, dev, True)
        self.x1 = D.f64(n, dev, True); self.v1 = D.f64(n, dev, True)
        self.E0 = D.f64(g, dev, True); self.Es = D.f64(g, dev, True); self.Fs = D.f64(g, dev, True)
        self.E1 = D.f64(g, dev, True); self.j0 = D.f64(g, dev, True)
        self.acc = D.f64(2 * g + (4 * g if self.det else 0), dev, True)     # det: + int64[4g] fixed-point words
        # [r, mean j1, EE, iterations | residual of every iteration of the step]
        self.stats = D.f64(4 + self.maxiter, dev, True)
        # enqueue-ahead Picard loop (see SheathSim.picard): the iterations the previous step needed are
        # queued without a host round trip each, guarded by a device flag the field kernel raises
        self.enqueue_ahead = os.environ.get("PIC_ENQUEUE_AHEAD", "1") != "0"   # env: A/B runs
        self.ctl = torch.zeros(1, dtype=torch.int32, device=dev)
        self._prev_hist = None
        self.range_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.last_iters, self.last_resid = 0, 1.0
        self.kernel_launches = 0
        self.iter_events = None     # set to a list to record a CUDA-event pair per particle-kernel launch
        self.light_iterations = self.qm_arrays is None
        self.x1b = None              # second n+1 position buffer (allocated with the first light push)
        self.Fs_prev = None
        self._ratio, self._r1 = None, None
        self.j1_repairs = 0
        self._sort_params = None
        self._perm2 = None

    def upload(self, x0, v0, E0=None):
        s = slice(self.start, self.stop)
        self.x0[:self.N].copy_(torch.as_tensor(np.ascontiguousarray(x0[s])))
        self.v0[:self.N].copy_(torch.as_tensor(np.ascontiguousarray(v0[s])))
        if E0 is not None:
            self.E0.copy_(torch.as_tensor(np.ascontiguousarray(E0)))

    def _expect_last(self, k, hist):
        """Will iteration k (1-based) be the last one?  (Same predictor as SheathSim.)"""
        if k >= self.maxiter or self._ratio is None:
            return True
        pred = self._r1 if k == 1 else hist[-1] * self._ratio
        return pred is None or pred <= 100.0 * self.tol

    def sort_by_cell(self):
        """Counting sort of the store by cell (pic_dev_dd_sort_by_cell) into the Picard scratch
        arrays; the original index of every particle rides along as a payload so that
        download() can return the arrays in the caller's order."""
        n = max(self.N, 1)
        if self.det:
            # reproducible build: stable LSD radix sort (equal cells keep their previous order), int32 payload
            first = self._sort_params is None
            if first:
                self._sort_scratch = torch.zeros(D.sort_stable_scratch_size(n), dtype=torch.int32, device=self.dev)
                self._sort_params = _lib.DDParams(self.N, self.N, self.Ng, 128, self.dx, self.dt, self.L, self.p2c,
                                                  (C.c_double * 2)(0., 0.), (C.c_double * 2)(1., 1.))
                if self.track_order:
                    self.perm = torch.empty(n, dtype=torch.int32, device=self.dev)
                    self._perm2 = torch.empty(n, dtype=torch.int32, device=self.dev)
            where = C.c_int(0)
            if self.track_order:
                _lib.call("pic_dev_dd_sort_by_cell_stable2", C.byref(self._sort_params), D.ptr(self.x0), D.ptr(self.v0),
                          D.ptr(self.x1), D.ptr(self.v1), D.ptr(self.perm), D.ptr(self._perm2), 1 if first else 0,
                          D.ptr(self._sort_scratch), self._sort_scratch.numel(), C.byref(where), D.stream())
            else:
                _lib.call("pic_dev_dd_sort_by_cell_stable", C.byref(self._sort_params), D.ptr(self.x0), D.ptr(self.v0),
                          D.ptr(self.x1), D.ptr(self.v1), D.ptr(self._sort_scratch), self._sort_scratch.numel(),
                          C.byref(where), D.stream())
            self.kernel_launches += 5 * ((max(1, (self.Ng - 1).bit_length()) + 7) // 8)
            if where.value:
                self.x0, self.x1 = self.x1, self.x0
                self.v0, self.v1 = self.v1, self.v0
                if self.track_order:
                    self.perm, self._perm2 = self._perm2, self.perm
            return
        if self._sort_params is None:
            if self.track_order:
                self.perm = torch.arange(n, dtype=torch.float64, device=self.dev)
                self._perm2 = torch.empty_like(self.perm)
            self._sort_counts = torch.zeros(D.sort_counts_size(self.Ng), dtype=torch.int32, device=self.dev)
            self._sort_params = _lib.DDParams(self.N, self.N, self.Ng, 0, self.dx, self.dt, self.L, self.p2c,
                                              (C.c_double * 2)(0., 0.), (C.c_double * 2)(1., 1.))
        _lib.call("pic_dev_dd_sort_by_cell", C.byref(self._sort_params), D.ptr(self.x0), D.ptr(self.v0),
                  D.ptr(self.perm), None, D.ptr(self.x1), D.ptr(self.v1), D.ptr(self._perm2) if self.track_order else None, None,
                  D.ptr(self._sort_counts), D.stream())
        self._sort_params.flags = 32          # later sorts see a nearly sorted store: global-cursor path
        self.kernel_launches += 3
        self.x0, self.x1 = self.x1, self.x0
        self.v0, self.v1 = self.v1, self.v0
        if self.track_order:
            self.perm, self._perm2 = self._perm2, self.perm

    def _allreduce_acc(self):
        """Sum over ranks of the deposits; reproducible build: of the fixed-point words (exact in any order)."""
        if self.det:
            if self.comm.enabled and self.comm.world > 1:
                self.comm.allreduce_sum(self.acc[2 * self.Ng:].view(torch.int64))
        else:
            self.comm.allreduce_sum(self.acc)

    def push(self):
        """particle_push_p: Picard loop + commit (x wrapped into [0,L)).  Returns (k, r)."""
        st = D.stream()
        P = C.byref(self.params)
        if self.sort_every and self.t % self.sort_every == 0:
            self.sort_by_cell()
        self.t += 1
        self.Es.copy_(self.E0)
        _lib.call("pic_dev_smooth", D.ptr(self.Es), D.ptr(self.Fs), self.Ng, 0, st)
        if self.stats.numel() < 4 + self.maxiter:          # maxiter was raised after construction
            self.stats = D.f64(4 + self.maxiter, self.dev, True)
        self.stats.zero_()
        self.ctl.zero_()
        base_flags = self.params.flags & ~8
        light_ok = self.light_iterations
        if light_ok and self.x1b is None:
            self.x1b = D.f64(max(self.N, 1), self.dev, True)
            self.Fs_prev = D.f64(self.Ng, self.dev, True)
        # the n+1 positions ping-pong between two buffers so that the inputs of the last iteration
        # survive it (needed by the repair pass when that iteration turns out to have been light)
        pairs = [(self.x1, self.x1b), (self.x1b, self.x1)] if light_ok else [(self.x1, self.x1)]
        rhist = D.ptr(self.stats) + 4 * 8
        queued = []                      # per iteration launched: (full, events or None)
        hist = []

        def launch(j, full):
            # v1 and j1 (the current at n+1) are only used after the loop: iterations that are not expected
            # to be the last one run "light" (flags bit3: no v1 store, no j1 lookup / deposit: 32 instead of
            # 40 B/particle); the prediction comes from the regular contraction of the residual, a wrong one
            # costs the repair pass below
            xin, xout = pairs[j % len(pairs)]
            self.params.flags = base_flags | (0 if full else 8)
            ev = None
            if self.iter_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            if self.qm_arrays is not None:
                _lib.call("pic_dev_pypic_picard_iter_qm", P, D.ptr(self.x0), D.ptr(self.v0), D.ptr(xin), D.ptr(xout),
                          D.ptr(self.v1), D.ptr(self.qm_arrays[0]), D.ptr(self.qm_arrays[1]), D.ptr(self.Fs), D.ptr(self.acc),
                          1 if j == 0 else 0, D.ptr(self.range_err), D.ptr(self.ctl), st)
            else:
                _lib.call("pic_dev_pypic_picard_iter3", P, D.ptr(self.x0), D.ptr(self.v0), D.ptr(xin), D.ptr(xout),
                          D.ptr(self.v1), D.ptr(self.Fs), D.ptr(self.acc), 1 if j == 0 else 0, D.ptr(self.range_err),
                          D.ptr(self.ctl), st)
            if ev is not None:
                ev[1].record()
            self._allreduce_acc()
            _lib.call("pic_dev_pypic_field_update2", P, D.ptr(self.acc), D.ptr(self.E0), D.ptr(self.Es), D.ptr(self.Fs),
                      D.ptr(self.E1), D.ptr(self.j0), D.ptr(self.stats), D.ptr(self.Fs_prev) if light_ok else None,
                      rhist, D.ptr(self.ctl), self.tol, self.maxiter, st)
            queued.append((full, ev))

        def outcome():
            sv = D.read_f64(self.stats, 4 + self.maxiter)
            kk = int(sv[3])
            return kk, [float(v) for v in sv[4:4 + kk]]

        k = 0
        ahead = self._prev_hist if (self.enqueue_ahead and self.maxiter >= 1) else None
        if ahead:
            for j in range(min(len(ahead), self.maxiter)):
                launch(j, (not light_ok) or self._expect_last(j + 1, ahead[:j]))
            k, hist = outcome()
        r = hist[-1] if hist else 1.0
        while (r > self.tol) and (k < self.maxiter) and k == len(queued):
            launch(k, (not light_ok) or self._expect_last(k + 1, hist))
            k, hist = outcome()
            r = hist[-1]
        if self.iter_events is not None:
            self.iter_events.extend(ev for _, ev in queued[:k])       # launches behind the end of the loop were no-ops
        self.kernel_launches += 2 * k
        full = queued[k - 1][0] if k > 0 else True
        last_in, last_out = pairs[(k - 1) % len(pairs)] if k > 0 else pairs[0]
        self._prev_hist = list(hist) if hist else self._prev_hist
        self.params.flags = base_flags
        if k > 0 and not full:
            _lib.call("pic_dev_pypic_j1_repair", P, D.ptr(self.x0), D.ptr(self.v0), D.ptr(last_in), D.ptr(last_out),
                      D.ptr(self.Fs_prev), D.ptr(self.v1), 1 if k == 1 else 0, D.ptr(self.acc), D.ptr(self.range_err), st)
            self._allreduce_acc()
            _lib.call("pic_dev_pypic_j1_finish", P, D.ptr(self.acc), D.ptr(self.j0), D.ptr(self.stats), st)
            self.kernel_launches += 2
            self.j1_repairs += 1
        if hist:
            ratios = [b_ / a_ for a_, b_ in zip(hist, hist[1:]) if a_ > 0.0]
            self._r1 = hist[0]
            if ratios:
                self._ratio = max(ratios)
        if k > 0:
            if light_ok:
                spare = self.x0
                self.x0 = last_out
                self.x1, self.x1b = (last_in, spare) if last_in is not last_out else (spare, self.x1b)
            else:
                self.x0, self.x1 = self.x1, self.x0
            self.v0, self.v1 = self.v1, self.v0
            self.E0, self.E1 = self.E1, self.E0
            # x1 = x1 % L (pypic.py:277) is applied lazily: on load by the next push, or by download()
        self.last_iters, self.last_resid = k, r
        return k, r

    def slot_of(self, i):
        """Store slot of the particle the caller uploaded as number i (local index)."""
        if self.perm is None:
            return int(i)
        return int(torch.nonzero(self.perm[:self.N] == float(i))[0, 0].item())

    def diagnostics(self, m=me):
        s = D.read_f64(self.stats, 4)
        if self.det:              # fixed-order reduction
            sc2 = D.f64(2, self.dev, True)
            _lib.call("pic_dev_moments", D.ptr(self.v0), self.N, D.ptr(sc2), D.stream())
            self.comm.allreduce_sum(sc2)
            return dict(EE=float(s[2]), KE=self.p2c_raw * (m / 2. * float(D.read_f64(sc2, 2)[1])), jbias=float(s[1]))
        sc = D.f64(1, self.dev, True)
        _lib.call("pic_dev_sum_sq", D.ptr(self.v0), self.N, m / 2., D.ptr(sc), D.stream())
        self.comm.allreduce_sum(sc)
        # pypic.py:571-574: EE = sum(eps0 E^2 dx/2), KE = p2c*sum(me v^2/2), j_bias = mean(j0)
        return dict(EE=float(s[2]), KE=self.p2c_raw * float(sc.item()), jbias=float(s[1]))

    def download(self):
        x, v = self.x0[:self.N].cpu().numpy() % self.L, self.v0[:self.N].cpu().numpy()
        if self.perm is not None:            # undo the cell sort: slot s holds original particle perm[s]
            idx = self.perm[:self.N].cpu().numpy().astype(np.int64)
            xo, vo = np.empty_like(x), np.empty_like(v)
            xo[idx] = x; vo[idx] = v
            x, v = xo, vo
        return dict(x0=x, v0=v, E0=self.E0.cpu().numpy(), j0=self.j0.cpu().numpy())

    def check(self):
        D.check_range(self.range_err, "pypic push")


class ExplicitSim:
    """PIC_L.main's loop: rho -> Poisson -> E -> kick-drift-kick -> wrap, with the deposit of
    the NEXT step fused into the push kernel."""

    def __init__(self, N, Ng, dx, dt, p2c, q=(-e, -e), m=(me, me), n_split=None, deposit="warp", comm=None,
                 device=None, sort_every=0, track_order=True):
        self.dev = D.require_cuda(device)
        self.comm = comm if comm is not None else Comm()
        self.N_global = int(N)
        self.start, self.stop = shard_range(N, self.comm.rank, self.comm.world)
        self.N = self.stop - self.start
        ns = int(N) if n_split is None else int(n_split)
        self.n_split = local_split(ns, self.start, self.stop)
        self.Ng, self.dx, self.dt, self.p2c = int(Ng), float(dx), float(dt), float(p2c)
        self.L = dx * (Ng - 1)              # PIC_L.py:645; the wrap length is L+dx
        if sort_every and deposit == "warp":
            deposit = "window"               # see PeriodicImplicitSim
        # "window-det": the REPRODUCIBLE build of the window kernel (flags bit7, as SheathSim's): every addition to the
        # global accumulator is an integer addition on fixed-point words, the cell sort is the stable radix sort
        # (with the original index as its payload when track_order), the kinetic energy is summed in a fixed
        # order -- two runs give bit-identical output
        flags = {"window": 0, "window-big": 16, "warp": 4, "atomic": 1 | 4, "window-det": 128}[deposit]
        self.det = deposit == "window-det"
        self.sort_every = int(sort_every)
        self.t = 0
        self.perm = None
        self.track_order = bool(track_order)
        self._sort_params = None
        self._perm2 = None
        self.params = _lib.LParams(self.N, self.n_split, self.Ng, flags, self.dx, self.dt, self.L, self.p2c,
                                   (C.c_double * 2)(*q), (C.c_double * 2)(*m))
        dev, n, g = self.dev, max(self.N, 1), self.Ng + 1
        self.x = D.f64(n, dev, True); self.v = D.f64(n, dev, True)
        self.q_arr = None
        self.rho_acc = D.f64(3 * g if self.det else g, dev, True)       # det: + int64[2g] fixed-point words
        self.rho = D.f64(g, dev, True); self.phi = D.f64(g, dev, True); self.E = D.f64(g, dev, True)
        self.work = D.f64(13 * g, dev, True)
        self.stats = D.f64(4, dev, True)
        self.range_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.kernel_launches = 0
        self._have_rho = False
        self.q, self.m = tuple(q), tuple(m)

    def upload(self, x, v):
        s = slice(self.start, self.stop)
        self.x[:self.N].copy_(torch.as_tensor(np.ascontiguousarray(x[s])))
        self.v[:self.N].copy_(torch.as_tensor(np.ascontiguousarray(v[s])))
        self._have_rho = False

    def _deposit_initial(self):
        """rho of the current positions (PIC_L.py:763) -- only needed before the first step."""
        if self.det:
            self.rho_acc.zero_()
            _lib.call("pic_dev_l_deposit_fixed", C.byref(self.params), D.ptr(self.x), D.ptr(self.rho_acc), D.ptr(self.range_err),
                      D.stream())
            self.kernel_launches += 1
            return
        qarr = torch.empty(max(self.N, 1), dtype=torch.float64, device=self.dev)
        qarr[:self.n_split] = self.q[0]
        qarr[self.n_split:] = self.q[1]
        tmp = D.f64(self.Ng + 1, self.dev, True)
        _lib.call("pic_dev_l_weight", D.ptr(self.x), D.ptr(qarr), None, D.ptr(tmp), self.N, self.Ng, self.dx, self.p2c,
                  D.ptr(self.range_err), D.stream())
        # l_weight already folded the ends; un-fold so that field_solve's fold is a no-op:
        # field_solve folds rho_acc[0]+rho_acc[-1]; give it the folded value split as (v, 0)
        self.rho_acc.copy_(tmp)
        self.rho_acc[-1] = 0.0
        self.kernel_launches += 2

    def field_solve(self):
        """PIC_L.py:763-766: (folded) rho -> phi (periodic Poisson, -max) -> E."""
        if not self._have_rho:
            self._deposit_initial()
        if self.det and self.comm.enabled and self.comm.world > 1:
            self.comm.allreduce_sum(self.rho_acc[self.Ng + 1:].view(torch.int64))      # exact in any order
        else:
            self.comm.allreduce_sum(self.rho_acc)
        _lib.call("pic_dev_l_field_solve", C.byref(self.params), D.ptr(self.rho_acc), D.ptr(self.rho), D.ptr(self.phi),
                  D.ptr(self.E), D.ptr(self.work), D.ptr(self.stats), D.stream())
        self.kernel_launches += 3

    def push(self):
        """PIC_L.py:767-768 + the deposit of the next step's rho."""
        _lib.call("pic_dev_l_push_deposit", C.byref(self.params), D.ptr(self.x), D.ptr(self.v), D.ptr(self.E),
                  D.ptr(self.rho_acc), D.ptr(self.range_err), D.stream())
        self.kernel_launches += 1
        self._have_rho = True

    def _sort_stable(self):
        """Reproducible build: LSD radix sort by (species, cell); equal cells keep their previous order."""
        n = max(self.N, 1)
        first = self._sort_params is None
        if first:
            self._x2 = torch.empty_like(self.x); self._v2 = torch.empty_like(self.v)
            self._sort_scratch = torch.zeros(D.sort_stable_scratch_size(n), dtype=torch.int32, device=self.dev)
            self._sort_params = _lib.DDParams(self.N, self.n_split, self.Ng, 128, self.dx, self.dt, self.L, self.p2c,
                                              (C.c_double * 2)(0., 0.), (C.c_double * 2)(1., 1.))
            if self.track_order:
                self.perm = torch.empty(n, dtype=torch.int32, device=self.dev)
                self._perm2 = torch.empty(n, dtype=torch.int32, device=self.dev)
        where = C.c_int(0)
        if self.track_order:
            _lib.call("pic_dev_dd_sort_by_cell_stable2", C.byref(self._sort_params), D.ptr(self.x), D.ptr(self.v), D.ptr(self._x2),
                      D.ptr(self._v2), D.ptr(self.perm), D.ptr(self._perm2), 1 if first else 0, D.ptr(self._sort_scratch),
                      self._sort_scratch.numel(), C.byref(where), D.stream())
        else:
            _lib.call("pic_dev_dd_sort_by_cell_stable", C.byref(self._sort_params), D.ptr(self.x), D.ptr(self.v), D.ptr(self._x2),
                      D.ptr(self._v2), D.ptr(self._sort_scratch), self._sort_scratch.numel(), C.byref(where), D.stream())
        passes = (max(1, (self.Ng - 1).bit_length()) + 7) // 8
        self.kernel_launches += 5 * passes * ((self.n_split > 0) + (self.n_split < self.N))
        if where.value:
            self.x, self._x2 = self._x2, self.x
            self.v, self._v2 = self._v2, self.v
            if self.track_order:
                self.perm, self._perm2 = self._perm2, self.perm

    def sort_by_cell(self):
        """Counting sort by (species, cell) with the original index as a payload."""
        if self.det:
            return self._sort_stable()
        n = max(self.N, 1)
        if self._sort_params is None:
            if self.track_order:
                self.perm = torch.arange(n, dtype=torch.float64, device=self.dev)
                self._perm2 = torch.empty_like(self.perm)
            self._x2 = torch.empty_like(self.x); self._v2 = torch.empty_like(self.v)
            self._sort_counts = torch.zeros(D.sort_counts_size(self.Ng), dtype=torch.int32, device=self.dev)
            self._sort_params = _lib.DDParams(self.N, self.n_split, self.Ng, 0, self.dx, self.dt, self.L, self.p2c,
                                              (C.c_double * 2)(0., 0.), (C.c_double * 2)(1., 1.))
        _lib.call("pic_dev_dd_sort_by_cell", C.byref(self._sort_params), D.ptr(self.x), D.ptr(self.v),
                  D.ptr(self.perm), None, D.ptr(self._x2), D.ptr(self._v2), D.ptr(self._perm2) if self.track_order else None, None,
                  D.ptr(self._sort_counts), D.stream())
        self._sort_params.flags = 32          # later sorts see a nearly sorted store: global-cursor path
        self.kernel_launches += 3
        self.x, self._x2 = self._x2, self.x
        self.v, self._v2 = self._v2, self.v
        if self.track_order:
            self.perm, self._perm2 = self._perm2, self.perm

    def step(self):
        if self.sort_every and self.t % self.sort_every == 0:
            self.sort_by_cell()
        self.t += 1
        self.field_solve()
        self.push()

    def field_energy(self):
        return float(D.read_f64(self.stats, 1)[0])

    def kinetic_energy(self, m=me):
        """sum(m v^2 / 2) over all particles (PIC_L.py:698 uses me for every particle)."""
        if self.det:          # fixed-order reduction
            sc2 = D.f64(2, self.dev, True)
            _lib.call("pic_dev_moments", D.ptr(self.v), self.N, D.ptr(sc2), D.stream())
            self.kernel_launches += 1
            self.comm.allreduce_sum(sc2)
            return m / 2. * float(D.read_f64(sc2, 2)[1])
        sc = D.f64(1, self.dev, True)
        _lib.call("pic_dev_sum_sq", D.ptr(self.v), self.N, m / 2., D.ptr(sc), D.stream())
        self.kernel_launches += 1
        self.comm.allreduce_sum(sc)
        return float(D.read_f64(sc, 1)[0])

    def download(self):
        x, v = self.x[:self.N].cpu().numpy(), self.v[:self.N].cpu().numpy()
        if self.perm is not None:
            idx = self.perm[:self.N].cpu().numpy().astype(np.int64)
            xo, vo = np.empty_like(x), np.empty_like(v)
            xo[idx] = x; vo[idx] = v
            x, v = xo, vo
        return dict(x=x, v=v, rho=self.rho.cpu().numpy(), phi=self.phi.cpu().numpy(), E=self.E.cpu().numpy())

    def check(self):
        D.check_range(self.range_err, "PIC_L step")
