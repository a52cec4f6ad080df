"""Device-resident driver of PIC_L_DD.py's bounded two-species implicit sheath
(PIC_L_DD.main_i, PIC_L_DD.py:415-551) on top of the C ABI.

State in HBM (structure of arrays, fp64): x0,u0 (time n), x1,u1 (Picard scratch /
time n+1; the commit is a pointer swap), optional passive v0,w0, int8 active flags,
and the small grid arrays.  One Picard iteration = one fused particle kernel
(gather+push+absorb+deposit jh,j1) + [one all-reduce if sharded] + one field kernel.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, device as D
from .dist import Comm, local_split, shard_range
from .rng import LegacyDraws, sheath_reinject_draws

epsilon0 = 8.854E-12
e = 1.602E-19
mp = 1.67E-27
me = 9.11E-31
kb = 1.38E-23


class SheathSim:
    def __init__(self, N, Ng, dx, dt, p2c, q=(-e, e), m=(me, mp), n_split=None, tol=1e-5, maxiter=20,
                 kBT=(None, None), gamma=0.0, carry_vw=True, deposit="window", tiles="smem",
                 rng="host", seed=1, draws=None, comm=None, device=None, sort_every=0, elide_u=True,
                 enqueue_ahead=True, reduce="nccl", track_order=None, vion_after=None):
        self.dev = D.require_cuda(device)
        self.comm = comm if comm is not None else Comm()
        self.N_global = int(N)
        n_split_g = int(N) // 2 if n_split is None else int(n_split)
        self.start, self.stop = shard_range(N, self.comm.rank, self.comm.world)
        self.N = self.stop - self.start
        self.n_split = local_split(n_split_g, self.start, self.stop)
        self._n_split_g = n_split_g
        self.Ng, self.dx, self.dt, self.p2c = int(Ng), float(dx), float(dt), float(p2c)
        self.L = dx * (Ng - 1)
        self.tol, self.maxiter, self.gamma = float(tol), int(maxiter), float(gamma)
        self.q, self.m = tuple(q), tuple(m)
        self.kBT = kBT
        self.carry_vw = bool(carry_vw)
        self.rng_mode = rng
        self.seed = int(seed)
        self.draws = draws if draws is not None else LegacyDraws()
        # sort_every=None: the interval the window kernel is tuned for -- 12 steps with the 15-node deposit windows
        # (grids up to 4352 nodes), 8 with the 7-node ones (profiles/r2_window_width_sheath.txt)
        if sort_every is None:
            sort_every = 12 if int(Ng) <= 4352 else 8
        self.sort_every = int(sort_every)
        # track_order: a cell-sorted store that keeps the reference's particle NUMBERING -- the sort
        # carries the original index of every particle as a payload (self.oid), the passive v0,w0 stay
        # in original order (they are never streamed), re-injection draws are made in original-index
        # order (PIC_L_DD.py:429-450) and download() returns the caller's order.  Default: on whenever a
        # sorted run needs it (host RNG parity or carried v,w); bench mode (Philox, no v,w) runs without.
        self.track = (bool(track_order) if track_order is not None
                      else (self.sort_every > 0 and (rng == "host" or self.carry_vw)))
        # vionout (PIC_L_DD.py:497-503): +/-u0 of electrons absorbed in steps t > vion_after
        self.vion_after = None if vion_after is None else int(vion_after)
        self._vion = []                 # (step, iteration, global index, value)
        # deposit: "window" = TMA-staged private-window kernel over contiguous chunks (default),
        # "window-ldg" = the same with register-prefetched loads instead of the TMA ring,
        # "atomic" = one shared-memory atomicAdd per contribution, "warp" = grid-stride kernel
        # with warp-uniform pre-reduction; tiles: "smem" or "global" (grid too large for smem)
        # "window-big" forces the large-grid build of the window kernel (per-warp field windows, no
        # whole-grid tile; chosen automatically when the grid does not fit shared memory)
        # "window-blocked": static round-robin of whole chunks instead of dynamically scheduled slices (bit6)
        # "window-det": the REPRODUCIBLE build of the default kernel (bit7): every addition to the global
        # accumulators is an integer addition on 2x64-bit fixed-point words (order-independent, also
        # across ranks: the all-reduce is an int64 sum) and the cell sort is the stable radix sort, so
        # two runs -- and runs on different numbers of CTAs -- give bit-identical states
        flags = {"window": 0, "window-blocked": 64, "window-big": 16, "window-ldg": 8, "atomic": 1,
                 "warp": 4, "window-det": 128}[deposit] | (2 if tiles == "global" else 0)
        self.det = deposit == "window-det"
        # (with the tracked order the stable sort carries the original-index payload: pic_dev_dd_sort_by_cell_stable2)
        self.params = _lib.DDParams(self.N, self.n_split, self.Ng, flags, self.dx, self.dt, self.L, self.p2c,
                                    (C.c_double * 2)(*self.q), (C.c_double * 2)(*self.m))
        dev = self.dev
        n = max(self.N, 1)
        self.x0 = D.f64(n, dev, True); self.u0 = D.f64(n, dev, True)
        self.x1 = D.f64(n, dev, True); self.u1 = D.f64(n, dev, True)
        # second n+1 position buffer: the Picard iterations ping-pong between x1 and x1b so that the
        # input of the last iteration survives it (needed if its velocities must be repaired)
        self.elide_u = bool(elide_u)
        self.x1b = D.f64(n, dev, True) if self.elide_u else None
        self._ratio, self._r1 = None, None      # residual contraction observed so far
        self.u_repairs = 0
        self.v0 = D.f64(n, dev, True) if carry_vw else None
        self.w0 = D.f64(n, dev, True) if carry_vw else None
        self.active = torch.ones(n, dtype=torch.int8, device=dev)
        g = self.Ng
        self.E0 = D.f64(g, dev, True); self.Es = D.f64(g, dev, True); self.E1 = D.f64(g, dev, True)
        self.Es_prev = D.f64(g, dev, True)
        self.j0 = D.f64(g, dev, True)
        # reproducible build: [jh | j1 | 4 counts] as fp64 followed by the int64 words [hi(2g) | lo(2g)]
        self.acc = D.f64(2 * g + 4 + (4 * g if self.det else 0), dev, True)
        # reduce: how a sharded run sums the accumulators over ranks -- "nccl": one all-reduce per deposit
        # (default); "p2p": the field kernel reads every rank's accumulators over NVLink peer memory
        # itself (pypic_b200/p2p.py; the particle kernels then add to the shared buffer and self.acc
        # receives the sum)
        self.p2p = None
        if reduce == "p2p" and self.comm.world > 1:
            if self.det or g > 32768:
                raise ValueError("reduce='p2p' serves the default build with Ng <= 32768")
            from .p2p import PeerAccumulators
            self.p2p = PeerAccumulators(self.comm, 2 * g + 4, dev)
        elif reduce not in ("nccl", "p2p"):
            raise ValueError("reduce must be 'nccl' or 'p2p'")
        self.wall_cum = D.f64(4, dev, True)
        # [r, mean j1, EE, iterations | 4 doubles of reduction scratch | residual of every iteration of the step]
        # ... | sum u0, sum u0^2 (diagnostics pass, or the step's first iteration when fused_moments is on) |
        # correction of the re-injection: sum(u_new - u_old), sum(u_new^2 - u_old^2)]
        self.stats = D.f64(8 + self.maxiter + 4, dev, True)
        # fused_moments: the first Picard iteration of every step also sums u0 and u0^2 (it streams u0 anyway) and
        # the re-injection kernel records what it changed, so np.std(u0) / KE of the state BEFORE the step
        # (PIC_L_DD.py:417, :549 of the previous step) come with the step's outcome read instead of a pass of
        # their own (pre_step_moments)
        self._fused_moments = False
        self._pre_mom = None
        # enqueue-ahead Picard loop: the iterations the previous step needed are queued back to back,
        # guarded by a device flag that the field kernel raises when the loop condition fails; the
        # host reads the outcome once per step instead of once per iteration
        self.enqueue_ahead = bool(enqueue_ahead) and os.environ.get("PIC_ENQUEUE_AHEAD", "1") != "0"   # env: A/B runs
        self.ctl = torch.zeros(1, dtype=torch.int32, device=dev)
        self._prev_hist = None
        self.range_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.dead_idx = torch.empty(n, dtype=torch.int32, device=dev)
        self.count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.block_counts = torch.zeros(2 * (n // 2048 + 2), dtype=torch.int64, device=dev)
        self.sort_counts = torch.zeros(D.sort_counts_size(g), dtype=torch.int32, device=dev) if self.sort_every else None
        self.sort_scratch = (torch.zeros(D.sort_stable_scratch_size(n), dtype=torch.int32, device=dev)
                             if self.sort_every and self.det else None)
        self.scalar = D.f64(2, dev, True)
        # absorption log (pic_dev_dd_picard_iter4): int32 [count,0,0,0 | {slot, original index, iteration, 0} x cap]
        self.dead_cap = int(min(max(n, 1), max(1 << 16, n // 64)))
        self.dead_buf = torch.zeros(4 + 4 * self.dead_cap, dtype=torch.int32, device=dev)
        self._log_valid = False         # the log describes the flags only after a step that started all-active
        self.dead_alt = torch.zeros_like(self.dead_buf)      # step() alternates the two logs (see _prologue)
        self.merge_prologue = True      # step(): re-injection + start-of-step clears in one launch
        self._begun = False
        self._log_guess = 256           # entries fetched with the counter in ONE read (grows with the counts seen)
        self._harvested = None
        self.oid = None                 # int32 original index per slot; None = identity (never sorted)
        self.oid_alt = None
        self._inv = None                # inverse of oid (built on demand for the host-draw thermostat)
        self.t = 0
        self._sorted_once = False
        self._sorts = 0
        self.heavy_sort_every = 1     # n > 1: only every n-th sort also re-sorts the heavy species (measured: no gain,
                                      # the kernel loses more on the ageing ion block than the sort saves)
        self.last_iters = 0
        self.last_resid = 1.0
        self.kernel_launches = 0
        self.vionout = []
        self.resid_trace = None     # set to a list to record the residual of every Picard iteration
        self.iter_events = None     # set to a list to record a CUDA-event pair per particle-kernel launch

    # ------------------------------------------------------------------ state I/O
    def upload(self, x0, u0, v0=None, w0=None, E0=None, active=None):
        """Global (unsharded) host arrays in, in the reference's particle order; each rank keeps its slice."""
        s = slice(self.start, self.stop)
        n = self.N                      # the arrays hold max(N, 1) slots so that an empty shard still has pointers
        self.x0[:n].copy_(torch.as_tensor(np.ascontiguousarray(x0[s])))
        self.u0[:n].copy_(torch.as_tensor(np.ascontiguousarray(u0[s])))
        if self.carry_vw:
            if v0 is not None:
                self.v0[:n].copy_(torch.as_tensor(np.ascontiguousarray(v0[s])))
            if w0 is not None:
                self.w0[:n].copy_(torch.as_tensor(np.ascontiguousarray(w0[s])))
        if E0 is not None:
            self.E0.copy_(torch.as_tensor(np.ascontiguousarray(E0)))
        if active is not None:
            self.active[:n].copy_(torch.as_tensor(np.ascontiguousarray(active[s]).astype(np.int8)))
        else:
            self.active.fill_(1)
        # the store is in the caller's order again: identity numbering, nothing logged, nothing sorted
        self.oid = self._inv = None
        self._log_valid = False
        self._harvested = None
        self._sorted_once = False

    def _to_original_order(self, t, kind):
        """A per-slot array of this shard in the reference's particle order (host copy)."""
        n = self.N
        if self.oid is None:
            return t[:n].cpu().numpy()
        out = torch.empty(max(n, 1), dtype=t.dtype, device=self.dev)
        _lib.call("pic_dev_scatter_" + kind, D.ptr(t), D.ptr(self.oid), D.ptr(out), n, D.stream())
        self.kernel_launches += 1
        return out[:n].cpu().numpy()

    def download(self):
        """The state of this shard in the reference's particle order (whatever the sort did)."""
        n = self.N
        act = self._to_original_order(self.active, "i8")
        out = dict(x0=self._to_original_order(self.x0, "f64"), u0=self._to_original_order(self.u0, "f64"),
                   active=act.astype(np.float64), E0=self.E0.cpu().numpy(), j0=self.j0.cpu().numpy())
        if self.carry_vw:
            # v,w are passive and live in original order.  A particle absorbed BEFORE the last Picard
            # iteration of the step holds the reference's zeros (PIC_L_DD.py:459-462 zero x1,u1,v1,w1 and
            # only active particles are pushed); one absorbed IN the last iteration keeps its values
            early = (act != 1) & (out["x0"] == 0.0)
            out["v0"] = np.where(early, 0.0, self.v0[:n].cpu().numpy())
            out["w0"] = np.where(early, 0.0, self.w0[:n].cpu().numpy())
        return out

    # ------------------------------------------------------------------ re-injection
    def _sigma(self, sp):
        kT = self.kBT[sp]
        return float(np.sqrt(kT / self.m[sp]))

    def _harvest_dead(self):
        """The slots absorbed during the last step, in ORIGINAL-index order: (slots, original local
        indices, Picard iteration each one died in or None).  From the absorption log when it is
        valid (a few hundred entries), otherwise by compacting the flag array.  Also tallies vionout
        (PIC_L_DD.py:497-503).  Cached until the next step."""
        if self._harvested is not None and self._harvested[0] == self.t:
            return self._harvested[1:]
        st = D.stream()
        slots = orig = iters = None
        if self._log_valid:
            # counter and (normally) all entries in ONE device->host read
            g = min(self._log_guess, self.dead_cap)
            buf = D.read_raw(self.dead_buf, 4 + 4 * g, np.int32)
            cnt = int(buf[0])
            if cnt <= self.dead_cap:
                if cnt > g:
                    buf = D.read_raw(self.dead_buf, 4 + 4 * cnt, np.int32)
                self._log_guess = max(256, 2 * cnt)
                ent = buf[4:4 + 4 * cnt].reshape(cnt, 4)
                slots, orig, iters = ent[:, 0].copy(), ent[:, 1].copy(), ent[:, 2].copy()
        if slots is None:
            _lib.call("pic_dev_compact_flags", D.ptr(self.active), self.N, 0, D.ptr(self.dead_idx), D.ptr(self.count),
                      D.ptr(self.block_counts), st)
            self.kernel_launches += 3
            cnt = int(D.read_raw(self.count, 1, np.int64)[0])
            slots = D.read_raw(self.dead_idx, cnt, np.int32) if cnt else np.zeros(0, dtype=np.int32)
            orig = slots
            if self.oid is not None and len(slots):
                dslots = D.to_dev(slots, self.dev, torch.int32)
                dorig = torch.empty(len(slots), dtype=torch.int32, device=self.dev)
                _lib.call("pic_dev_gather_i32", D.ptr(self.oid), D.ptr(dslots), D.ptr(dorig), len(slots), st)
                self.kernel_launches += 1
                orig = D.read_raw(dorig, len(slots), np.int32)
        order = np.argsort(orig, kind="stable")
        slots, orig = np.ascontiguousarray(slots[order]), np.ascontiguousarray(orig[order])
        iters = None if iters is None else iters[order]
        self._tally_vionout(slots, orig, iters)
        self._harvested = (self.t, slots, orig, iters)
        return slots, orig, iters

    def _tally_vionout(self, slots, orig, iters):
        """PIC_L_DD.py:497-503: for steps t > 2000 every absorbed electron (i < N/2) appends u0[i]
        (right wall) or -u0[i] (left wall), in (Picard iteration, index) order.  The deaths harvested at
        the start of step t happened in step t-1; u0 of that step is what the commit left in self.u1."""
        step = self.t - 1
        if self.vion_after is None or step <= self.vion_after or not len(slots):
            return
        sel = np.nonzero((orig.astype(np.int64) + self.start) < self.N_global // 2)[0]
        if not len(sel):
            return
        if iters is None:
            raise _lib.PicError(_lib.PIC_ERR_ARG, "vionout needs the absorption log of the step (the state was uploaded or "
                                "restored with absorbed particles, or the log overflowed)")
        ds = D.to_dev(slots[sel], self.dev, torch.int32)
        du = torch.empty(len(sel), dtype=torch.float64, device=self.dev)
        df = torch.empty(len(sel), dtype=torch.int8, device=self.dev)
        _lib.call("pic_dev_gather_f64", D.ptr(self.u1), D.ptr(ds), D.ptr(du), len(sel), D.stream())
        _lib.call("pic_dev_gather_i8", D.ptr(self.active), D.ptr(ds), D.ptr(df), len(sel), D.stream())
        self.kernel_launches += 2
        u = D.read_f64(du, len(sel)); f = D.read_raw(df, len(sel), np.int8)
        for k, j in enumerate(sel):
            self._vion.append((step, int(iters[j]), int(orig[j]) + self.start, float(u[k]) if f[k] == 0 else -float(u[k])))

    def collect_vionout(self):
        """The vionout list of the whole run (all ranks), in the reference's append order."""
        if self.vion_after is not None:
            self._harvest_dead()                  # the deaths of the last step have not been visited by a re-injection
        rows = list(self._vion)
        if self.comm.enabled and self.comm.world > 1:
            import torch.distributed as dist
            parts = [None] * self.comm.world
            dist.all_gather_object(parts, rows, group=self.comm.group)
            rows = [r for p in parts for r in p]
        rows.sort(key=lambda r: r[:3])
        self.vionout = [r[3] for r in rows]
        return self.vionout

    def _thermostat_host(self, dead_orig_global):
        """PIC_L_DD.py:419-427 with the legacy stream: one uniform per ACTIVE particle in index order;
        u < gamma redraws u,v,w from the ION temperature (kBTi for both species, as written).  Every
        rank runs the whole stream and applies the hits inside its own index range."""
        nd = len(dead_orig_global)
        n_active = self.N_global - nd
        hs = self.N_global // 2 if self._n_split_g is None else self._n_split_g
        k_split = hs - int(np.searchsorted(dead_orig_global, hs, side="left"))
        sig = [float(np.sqrt(self.kBT[1] / self.m[0])), float(np.sqrt(self.kBT[1] / self.m[1]))]
        hk, hu, hv, hw = self.draws.sheath_thermostat(n_active, k_split, self.gamma, sig[0], sig[1])
        if not len(hk):
            return 0
        # ordinal among the active particles -> global index: skip the dead slots below it
        gi = hk + np.searchsorted(dead_orig_global - np.arange(nd), hk, side="right")
        mine = (gi >= self.start) & (gi < self.stop)
        if not mine.any():
            return 0
        orig = (gi[mine] - self.start).astype(np.int32)
        if self.oid is not None:
            if self._inv is None:
                self._inv = torch.empty(max(self.N, 1), dtype=torch.int32, device=self.dev)
                _lib.call("pic_dev_invert_perm", D.ptr(self.oid), D.ptr(self._inv), self.N, D.stream())
            dorig = D.to_dev(orig, self.dev, torch.int32)
            dslot = torch.empty(len(orig), dtype=torch.int32, device=self.dev)
            _lib.call("pic_dev_gather_i32", D.ptr(self._inv), D.ptr(dorig), D.ptr(dslot), len(orig), D.stream())
        else:
            dorig = dslot = D.to_dev(orig, self.dev, torch.int32)
        dd = D.to_dev(np.stack([hu[mine], hv[mine], hw[mine]]), self.dev)
        _lib.call("pic_dev_dd_apply_draws3", D.ptr(dslot), D.ptr(dorig), None, D.ptr(dd[0]), D.ptr(dd[1]), D.ptr(dd[2]),
                  len(orig), None, D.ptr(self.u0), D.ptr(self.v0), D.ptr(self.w0), None, getattr(self, "_corr_ptr", None),
                  D.stream())
        self.kernel_launches += 3
        return int(mine.sum())

    def _prologue(self, philox, n_draws=0, base=0):
        """One launch for the re-injection and the start-of-step clears (pic_dev_dd_step_prologue).  It reads
        the log the previous step wrote and resets the OTHER log buffer, which the coming step then writes."""
        nst = 8 + self.maxiter
        if self.stats.numel() < nst + 4:
            self.stats = D.f64(nst + 4, self.dev, True)
        a = getattr(self, "_pro", None)
        if a is None:
            a = self._pro = _lib.DDPrologue()        # filled once; only what changes from step to step below
            a.log_cap = self.dead_cap
            a.sigma[0], a.sigma[1] = self._sigma(0), self._sigma(1)
            a.seed, a.global_offset = self.seed, self.start
            a.v0, a.w0 = D.ptr(self.v0), D.ptr(self.w0)
            a.wall_cum, a.ctl = D.ptr(self.wall_cum), D.ptr(self.ctl)
        a.log, a.next_log = self.dead_buf.data_ptr(), self.dead_alt.data_ptr()
        a.philox = 1 if philox else 0
        a.step = self.t
        a.x0, a.u0, a.active = self.x0.data_ptr(), self.u0.data_ptr(), self.active.data_ptr()
        a.orig = D.ptr(self.oid)
        a.n_draws = int(n_draws)
        if n_draws:
            a.slot, a.orig_of_draw = base + 32 * n_draws, base + 36 * n_draws
            a.xd, a.ud = base, base + 8 * n_draws
            a.vd, a.wd = (base + 16 * n_draws, base + 24 * n_draws) if self.carry_vw else (None, None)
        self._corr_ptr = self.stats.data_ptr() + 8 * (nst + 2) if (self.fused_moments and not philox) else None
        a.corr = self._corr_ptr
        a.Es, a.E0, a.stats = self.Es.data_ptr(), self.E0.data_ptr(), self.stats.data_ptr()
        a.nstats = nst + 2
        _lib.call("pic_dev_dd_step_prologue", C.byref(self.params), C.byref(a), D.stream())
        self.kernel_launches += 1
        self.dead_buf, self.dead_alt = self.dead_alt, self.dead_buf
        self._log_valid = True
        self._begun = True

    def reinject(self, merge=False):
        """PIC_L_DD.py:419-450: thermostat, then re-injection of the absorbed slots.  merge: the start-of-step
        clears of picard() ride in the same launch (step() asks for that)."""
        st = D.stream()
        if self.rng_mode == "philox":
            sig = (C.c_double * 2)(self._sigma(0), self._sigma(1))
            if self.gamma != 0.0:
                sgi = (C.c_double * 2)(float(np.sqrt(self.kBT[1] / self.m[0])), float(np.sqrt(self.kBT[1] / self.m[1])))
                _lib.call("pic_dev_dd_thermostat_philox", C.byref(self.params), D.ptr(self.u0), D.ptr(self.v0), D.ptr(self.w0),
                          D.ptr(self.active), D.ptr(self.oid), self.gamma, C.byref(sgi), self.seed, self.t, self.start, st)
                self.kernel_launches += 1
            if self.vion_after is not None:
                self._harvest_dead()
            # the slots named in the absorption log (a few hundred) are re-injected without touching the
            # flag array; the flag scan only runs when there is no valid log (first step, overflow).
            # With a tracked order the draws are keyed by the ORIGINAL index and v,w written there.
            if merge and self._log_valid:
                self._prologue(True)
                return None
            if self._log_valid:
                _lib.call("pic_dev_dd_reinject_philox_log", C.byref(self.params), D.ptr(self.dead_buf), self.dead_cap,
                          D.ptr(self.x0), D.ptr(self.u0), D.ptr(self.v0), D.ptr(self.w0), D.ptr(self.active),
                          D.ptr(self.oid), C.byref(sig), self.seed, self.t, self.start, st)
            _lib.call("pic_dev_dd_reinject_philox2", C.byref(self.params), D.ptr(self.x0), D.ptr(self.u0),
                      D.ptr(self.v0), D.ptr(self.w0), D.ptr(self.active), D.ptr(self.oid), C.byref(sig), self.seed, self.t,
                      self.start, D.ptr(self.dead_buf) if self._log_valid else None, self.dead_cap, st)
            self.kernel_launches += 1
            self._reset_log()
            return None
        merge = merge and self.gamma == 0.0      # the host thermostat adds its own share of the moments' correction first
        slots, orig, _ = self._harvest_dead()
        n_dead = len(slots)
        counts = self.comm.allgather_int(n_dead, device=self.dev)
        self._corr_ptr = None
        if self.fused_moments:
            self._corr_ptr = D.ptr(self.stats) + 8 * (8 + self.maxiter + 2)
            if not merge:
                _lib.call("pic_dev_zero", self._corr_ptr, 16, st)
        if self.gamma != 0.0:
            g_orig = orig.astype(np.int64) + self.start
            if self.comm.enabled and self.comm.world > 1:
                import torch.distributed as dist
                parts = [None] * self.comm.world
                dist.all_gather_object(parts, g_orig, group=self.comm.group)
                g_orig = np.concatenate(parts)
            self._thermostat_host(np.sort(g_orig))
        else:
            self.draws.sheath_thermostat_skip(self.N_global - sum(counts))
        sigma = np.where(orig >= self.n_split, self._sigma(1), self._sigma(0))
        xd, ud, vd, wd = sheath_reinject_draws(self.draws, counts, self.comm.rank, sigma, self.L)
        if self.gamma == 0.0:
            # the next step skips about as many thermostat uniforms: jump ahead while the GPU works
            self.draws.prefetch_skip(self.N_global - sum(counts))
        if n_dead:
            # one asynchronous copy from a pinned staging buffer: [x | u | v | w | (slot, orig) as int32]
            hs, ds = self._staging(n_dead)
            hv = hs.numpy()
            hv[0:n_dead] = xd; hv[n_dead:2 * n_dead] = ud; hv[2 * n_dead:3 * n_dead] = vd; hv[3 * n_dead:4 * n_dead] = wd
            iv = hv[4 * n_dead:5 * n_dead + 1].view(np.int32)
            iv[0:n_dead] = slots; iv[n_dead:2 * n_dead] = orig
            _lib.call("pic_dev_write", D.ptr(ds), D.ptr(hs), (5 * n_dead + 1) * 8, st)
            base = D.ptr(ds)
            if merge:
                self._prologue(False, n_dead, base)
                return n_dead
            _lib.call("pic_dev_dd_apply_draws3", base + 32 * n_dead, base + 36 * n_dead, base, base + 8 * n_dead,
                      base + 16 * n_dead if self.carry_vw else None, base + 24 * n_dead if self.carry_vw else None, n_dead,
                      D.ptr(self.x0), D.ptr(self.u0), D.ptr(self.v0), D.ptr(self.w0), D.ptr(self.active), self._corr_ptr, st)
            self.kernel_launches += 1
        elif merge:
            self._prologue(False)
            return 0
        self._reset_log()
        return n_dead

    def _staging(self, n):
        """(pinned host, device) staging pair for this step's draws.  Two pairs alternate: the copy of
        step t is long complete (the Picard loop of step t synchronises) when step t+2 refills its pair."""
        if not hasattr(self, "_stage") or self._stage[0][0].numel() < 5 * n + 2:
            cap = max(4096, 2 * (5 * n + 2))
            self._stage = [(torch.empty(cap, dtype=torch.float64, pin_memory=True),
                            torch.empty(cap, dtype=torch.float64, device=self.dev)) for _ in range(2)]
        return self._stage[self.t & 1]

    def _reset_log(self):
        """Every slot is alive now: the absorption log of the coming step starts empty."""
        _lib.call("pic_dev_zero", D.ptr(self.dead_buf), 16, D.stream())
        self._log_valid = True

    @property
    def fused_moments(self):
        return self._fused_moments

    @fused_moments.setter
    def fused_moments(self, on):
        # the fused sums are merged with fp64 atomics in scheduling order: not with the reproducible build, whose
        # diagnostics come from the fixed-order reduction of pic_dev_moments
        self._fused_moments = bool(on) and not self.det

    def sort_by_cell(self):
        """Counting sort by (species, cell) into the scratch arrays."""
        st = D.stream()
        if self.track:
            # the sort carries the original index of every particle (identity on the first sort); all
            # slots are alive here (re-injection ran before), so the flags need no permutation, and the
            # passive v,w are not moved at all
            P = self.params
            if self.det:
                # reproducible build: the STABLE radix sort with the payload -- the order inside a cell is the previous
                # order, so the private-window sums (and with them every bit of the run) do not depend on scheduling
                identity = self.oid is None
                if identity:
                    self.oid = torch.empty(max(self.N, 1), dtype=torch.int32, device=self.dev)
                if self.oid_alt is None:
                    self.oid_alt = torch.empty(max(self.N, 1), dtype=torch.int32, device=self.dev)
                where = C.c_int(0)
                _lib.call("pic_dev_dd_sort_by_cell_stable2", C.byref(P), D.ptr(self.x0), D.ptr(self.u0), D.ptr(self.x1),
                          D.ptr(self.u1), D.ptr(self.oid), D.ptr(self.oid_alt), 1 if identity else 0, D.ptr(self.sort_scratch),
                          self.sort_scratch.numel(), C.byref(where), st)
                passes = (max(1, (self.Ng - 1).bit_length()) + 7) // 8
                self.kernel_launches += 5 * passes * ((self.n_split > 0) + (self.n_split < self.N))
                if where.value:
                    self.x0, self.x1 = self.x1, self.x0
                    self.u0, self.u1 = self.u1, self.u0
                    self.oid, self.oid_alt = self.oid_alt, self.oid
                self._inv = None
                self._sorted_once = True
                self._sorts += 1
                return
            if self._sorted_once:
                P = _lib.DDParams(self.N, self.n_split, self.Ng, self.params.flags | 32, self.dx, self.dt, self.L, self.p2c,
                                  (C.c_double * 2)(*self.q), (C.c_double * 2)(*self.m))
            if self.oid_alt is None:
                self.oid_alt = torch.empty(max(self.N, 1), dtype=torch.int32, device=self.dev)
            _lib.call("pic_dev_dd_sort_by_cell2", C.byref(P), D.ptr(self.x0), D.ptr(self.u0), D.ptr(self.oid),
                      D.ptr(self.x1), D.ptr(self.u1), D.ptr(self.oid_alt), D.ptr(self.sort_counts), st)
            self.kernel_launches += 3
            self.oid, self.oid_alt = self.oid_alt, (self.oid if self.oid is not None else None)
            self._inv = None
            self._sorted_once = True
            self._sorts += 1
            self.x0, self.x1 = self.x1, self.x0
            self.u0, self.u1 = self.u1, self.u0
            return
        if self.det:
            # reproducible build: stable radix sort (the order inside a cell is the previous order)
            where = C.c_int(0)
            _lib.call("pic_dev_dd_sort_by_cell_stable", C.byref(self.params), D.ptr(self.x0), D.ptr(self.u0),
                      D.ptr(self.x1), D.ptr(self.u1), D.ptr(self.sort_scratch), self.sort_scratch.numel(),
                      C.byref(where), st)
            passes = (max(1, (self.Ng - 1).bit_length()) + 7) // 8
            self.kernel_launches += 5 * passes * ((self.n_split > 0) + (self.n_split < self.N))
            self._sorted_once = True
            self._sorts += 1
            if where.value:
                self.x0, self.x1 = self.x1, self.x0
                self.u0, self.u1 = self.u1, self.u0
            return
        # a store that was sorted a few steps ago is NEARLY sorted: the global-memory cursor path of
        # the sort (flags bit5) is then faster than the shared-memory one (2.5 vs 3.0 ms at 2e8
        # particles); the first sort of a random store takes the shared-memory path
        P = self.params
        ns = self.n_split
        light_only = (self._sorted_once and self._sorts % self.heavy_sort_every != 0 and 0 < ns < self.N
                      and ns % 2 == 0 and self.m[0] < self.m[1])
        self._sorts += 1
        if light_only:
            # the heavy species (ions) crosses a cell ~40x more slowly than the light one: between
            # its own (rarer) sorts only the light block [0, n_split) is sorted, into the scratch
            # arrays, and copied back -- 2.1 ms instead of 3.1 ms at 2e8 particles
            P = _lib.DDParams(ns, ns, self.Ng, self.params.flags | 32, self.dx, self.dt, self.L, self.p2c,
                              (C.c_double * 2)(*self.q), (C.c_double * 2)(*self.m))
            _lib.call("pic_dev_dd_sort_by_cell", C.byref(P), D.ptr(self.x0), D.ptr(self.u0), None, None,
                      D.ptr(self.x1), D.ptr(self.u1), None, None, D.ptr(self.sort_counts), st)
            _lib.call("pic_dev_copy", D.ptr(self.x0), D.ptr(self.x1), ns * 8, st)
            _lib.call("pic_dev_copy", D.ptr(self.u0), D.ptr(self.u1), ns * 8, st)
            self.kernel_launches += 5
            return
        if self._sorted_once:
            P = _lib.DDParams(self.N, self.n_split, self.Ng, self.params.flags | 32, self.dx, self.dt, self.L, self.p2c,
                              (C.c_double * 2)(*self.q), (C.c_double * 2)(*self.m))
        self._sorted_once = True
        _lib.call("pic_dev_dd_sort_by_cell", C.byref(P), D.ptr(self.x0), D.ptr(self.u0), None, None,
                  D.ptr(self.x1), D.ptr(self.u1), None, None, D.ptr(self.sort_counts), st)
        self.kernel_launches += 3
        self.x0, self.x1 = self.x1, self.x0
        self.u0, self.u1 = self.u1, self.u0

    # ------------------------------------------------------------------ one timestep
    def _expect_last(self, k, hist):
        """Will iteration k (1-based) be the last one?  The residual contracts by a very regular
        factor per iteration, so the next residual is predicted from the ratio observed so far; a
        wrong guess only costs the repair pass, never correctness."""
        if k >= self.maxiter or self._ratio is None:
            return True
        pred = self._r1 if k == 1 else hist[-1] * self._ratio
        return pred is None or pred <= 100.0 * self.tol

    def _acc_ptr(self):
        """Where the particle kernels deposit."""
        return self.p2p.mine if self.p2p is not None else D.ptr(self.acc)

    def _allreduce_acc(self):
        """One sum over ranks per deposit.  Reproducible build: the currents travel as int64
        fixed-point words (an integer all-reduce is exact in any order), the four absorbed counts
        as exact fp64 integers."""
        if self.p2p is not None:
            return self.acc                      # summed by the field kernel over peer memory
        if not self.det:
            return self.comm.allreduce_sum(self.acc)
        if self.comm.world > 1:
            g = self.Ng
            self.comm.allreduce_sum(self.acc[2 * g:2 * g + 4])
            self.comm.allreduce_sum(self.acc[2 * g + 4:].view(torch.int64))
        return self.acc

    def picard(self):
        """PIC_L_DD.py:452-545: Picard loop + commit.  Returns (iterations, residual).

        The velocities u1 and the current j1 at n+1 are only needed after the loop, so an iteration
        stores / deposits them only when it is expected to be the last one ("light" iterations
        stream 32 instead of 40 bytes per particle and skip a third of the deposit work); if the
        loop ends on a light iteration, pic_dev_dd_commit_u2 + pic_dev_dd_j1_finish recompute both
        from that iteration's inputs (kept intact by ping-ponging the position buffers).

        Enqueue-ahead: the residuals of consecutive steps differ by a few per cent, so the
        iterations the previous step needed are queued without waiting for their residuals; every
        launch is guarded by the device flag `ctl`, which the field kernel raises as soon as
        `r > tol and k < maxiter` fails, so a loop that ends earlier turns the rest of the queue into
        no-ops and one that needs more continues one iteration at a time."""
        st = D.stream()
        P = C.byref(self.params)
        if self.stats.numel() < 8 + self.maxiter + 4:      # maxiter was raised after construction
            self.stats = D.f64(8 + self.maxiter + 4, self.dev, True)
        nst = 8 + self.maxiter
        # clears the statistics and the moments; the re-injection correction behind them was written by this
        # step's reinject() and is cleared by the next one
        if self._begun:
            self._begun = False          # step()'s merged prologue did it
        else:
            _lib.call("pic_dev_dd_step_begin", D.ptr(self.Es), D.ptr(self.E0), self.Ng, D.ptr(self.wall_cum), D.ptr(self.stats),
                      nst + 2, D.ptr(self.ctl), st)
        mom_ptr = D.ptr(self.stats) + 8 * nst if self.fused_moments else None      # (never with the reproducible build)
        rhist = D.ptr(self.stats) + 8 * 8
        pairs = [(self.x1, self.x1b), (self.x1b, self.x1)] if self.elide_u else [(self.x1, self.x1)]
        queued = []                      # per iteration launched: (want_u, events or None)
        hist = []

        def launch(j, want_u):           # j: 0-based iteration
            xin, xout = pairs[j % len(pairs)]
            ev = None
            if self.iter_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            _lib.call("pic_dev_dd_picard_iter5", P, D.ptr(self.x0), D.ptr(self.u0), D.ptr(xin), D.ptr(xout),
                      D.ptr(self.u1) if want_u else None, D.ptr(self.active), D.ptr(self.Es), self._acc_ptr(),
                      1 if j == 0 else 0, D.ptr(self.range_err), D.ptr(self.ctl), D.ptr(self.dead_buf), self.dead_cap,
                      D.ptr(self.oid), j, mom_ptr if j == 0 else None, st)
            if j == 0 and mom_ptr is not None and self.comm.enabled and self.comm.world > 1:
                self.comm.allreduce_sum(self.stats[nst:nst + 4])        # one small collective per step (the diagnostics)
            if ev is not None:
                ev[1].record()
            if self.p2p is not None:
                pp = self.p2p
                _lib.call("pic_dev_dd_field_update_p2p", P, D.ptr(pp.peers_dev), pp.rank, pp.world, pp.next_seq(),
                          D.ptr(self.acc), D.ptr(self.wall_cum), D.ptr(self.E0), D.ptr(self.Es), D.ptr(self.E1),
                          D.ptr(self.j0), D.ptr(self.stats), D.ptr(self.Es_prev) if self.elide_u else None, rhist,
                          D.ptr(self.ctl), self.tol, self.maxiter, D.ptr(pp.err), st)
            else:
                self._allreduce_acc()
                _lib.call("pic_dev_dd_field_update2", P, D.ptr(self.acc), D.ptr(self.wall_cum), D.ptr(self.E0),
                          D.ptr(self.Es), D.ptr(self.E1), D.ptr(self.j0), D.ptr(self.stats),
                          D.ptr(self.Es_prev) if self.elide_u else None, rhist, D.ptr(self.ctl), self.tol, self.maxiter, st)
            queued.append((want_u, ev))

        def outcome():
            s = D.read_f64(self.stats, 8 + self.maxiter + 4)
            if self.fused_moments:
                # moments of the state BEFORE this step's re-injection / thermostat = after the previous step
                self._pre_mom = (float(s[nst] - s[nst + 2]), float(s[nst + 1] - s[nst + 3]))
            if self.p2p is not None:
                self.p2p.check()         # a timed-out wait summed incomplete accumulators: stop at once
            k = int(s[3])
            self._last_stats = [float(v) for v in s[:4]]
            return k, [float(v) for v in s[8:8 + k]]

        k = 0
        ahead = self._prev_hist if (self.enqueue_ahead and self.maxiter >= 1) else None
        if ahead:
            # the previous step's residuals stand in for this step's in the last-iteration predictor
            for j in range(min(len(ahead), self.maxiter)):
                launch(j, (not self.elide_u) or self._expect_last(j + 1, ahead[:j]))
            k, hist = outcome()
        r = hist[-1] if hist else 1.0
        while (r > self.tol) and (k < self.maxiter) and k == len(queued):
            launch(k, (not self.elide_u) or self._expect_last(k + 1, hist))
            k, hist = outcome()
            r = hist[-1]
        # launches behind the end of the loop were no-ops
        if self.iter_events is not None:
            self.iter_events.extend(ev + (wu, j == 0) for j, (wu, ev) in enumerate(queued[:k]))
        self.kernel_launches += 2 * k
        if self.resid_trace is not None:
            self.resid_trace.extend(hist)
        wrote_u = queued[k - 1][0] if k > 0 else True
        last_in, last_out = pairs[(k - 1) % len(pairs)] if k > 0 else pairs[0]
        if k > 0 and not wrote_u:
            # the loop ended on a light iteration: recompute its velocities and its j1
            _lib.call("pic_dev_dd_commit_u2", P, D.ptr(self.x0), D.ptr(self.u0), D.ptr(last_in), D.ptr(last_out),
                      D.ptr(self.active), D.ptr(self.Es_prev), D.ptr(self.u1), 1 if k == 1 else 0, self._acc_ptr(),
                      D.ptr(self.range_err), st)
            if self.p2p is not None:
                pp = self.p2p
                _lib.call("pic_dev_p2p_reduce", D.ptr(pp.peers_dev), pp.rank, pp.world, pp.next_seq(), pp.nacc,
                          D.ptr(self.acc), D.ptr(pp.err), st)
            self._allreduce_acc()
            _lib.call("pic_dev_dd_j1_finish", P, D.ptr(self.acc), D.ptr(self.wall_cum), D.ptr(self.j0), D.ptr(self.stats), st)
            self.kernel_launches += 2
            self.u_repairs += 1
        if hist:
            ratios = [b / a for a, b in zip(hist, hist[1:]) if a > 0.0]
            self._r1 = hist[0]
            if ratios:
                self._ratio = max(ratios)
            self._prev_hist = list(hist)
        # commit (PIC_L_DD.py:538-545): pointer swaps
        if k > 0:
            if self.elide_u:
                spare = self.x0
                self.x0 = last_out
                self.x1, self.x1b = (last_in, spare) if last_in is not last_out else (spare, self.x1b)
            else:
                self.x0, self.x1 = self.x1, self.x0
            self.u0, self.u1 = self.u1, self.u0
            self.E0, self.E1 = self.E1, self.E0
        self.last_iters, self.last_resid = k, r
        return k, r

    def step(self):
        self.reinject(merge=self.merge_prologue)
        if self.sort_every and self.t % self.sort_every == 0 and (self.track or (self.rng_mode == "philox" and not self.carry_vw)):
            self.sort_by_cell()
        out = self.picard()
        self.t += 1
        return out

    # ------------------------------------------------------------------ diagnostics
    def diagnostics(self):
        """EE, KE, jbias of PIC_L_DD.py:548-551 (KE uses me for every particle, as written)."""
        s = self._moments_and_stats()
        m1, m2 = float(s[-2]), float(s[-1])
        return dict(EE=float(s[2]), KE=me / 2. * m2, jbias=float(s[1]), kBTe=self.kBTe_from(m1, m2))

    def diagnostics_begin(self):
        """Launch the pass for the step's diagnostics without waiting for it; diagnostics_end() reads
        the numbers -- e.g. at the top of the next step, where the host has to wait for the device
        anyway (the values do not change in between: nothing touches u0 or the statistics)."""
        tail = self.stats[8 + self.maxiter:8 + self.maxiter + 2]
        _lib.call("pic_dev_moments", D.ptr(self.u0), self.N, D.ptr(tail), D.stream())
        self.kernel_launches += 1
        self.comm.allreduce_sum(tail)

    def pre_step_moments(self):
        """fused_moments: (sum u0, sum u0^2) over all ranks of the state BEFORE the step that just ran (i.e. after
        the step before it, re-injection and thermostat not yet applied) -- what the reference prints as
        np.std(u0) at the top of the step (PIC_L_DD.py:417) and sums into KE at the end of the previous one (:549)."""
        return self._pre_mom

    def step_stats(self):
        """EE and jbias of the step that just ran (PIC_L_DD.py:548,551) from the statistics the Picard loop's
        outcome read already fetched (no device access)."""
        return dict(EE=self._last_stats[2], jbias=self._last_stats[1])

    def diagnostics_end(self):
        s = D.read_f64(self.stats, 8 + self.maxiter + 2)
        m1, m2 = float(s[-2]), float(s[-1])
        return dict(EE=float(s[2]), KE=me / 2. * m2, jbias=float(s[1]), kBTe=self.kBTe_from(m1, m2))

    def _moments_and_stats(self):
        """One pass over the velocities for (sum u0, sum u0^2), summed over the ranks into the tail of
        the step's statistics, and ONE read of the whole block."""
        tail = self.stats[8 + self.maxiter:8 + self.maxiter + 2]
        _lib.call("pic_dev_moments", D.ptr(self.u0), self.N, D.ptr(tail), D.stream())
        self.kernel_launches += 1
        self.comm.allreduce_sum(tail)
        return D.read_f64(self.stats, 8 + self.maxiter + 2)

    def moments(self):
        """(sum u0, sum u0^2) over all ranks."""
        s = self._moments_and_stats()
        return float(s[-2]), float(s[-1])

    def kBTe_from(self, m1, m2):
        """np.std(u0)**2 * me / e (PIC_L_DD.py:417; dead slots count with their zeros, as there)."""
        n = float(self.N_global)
        return max(m2 / n - (m1 / n) ** 2, 0.0) * me / e

    def phi(self):
        """phih of the last Picard iteration: -cumtrapz(Eh) - max (PIC_L_DD.py:522-523);
        after the loop Es holds that Eh."""
        out = D.f64(self.Ng, self.dev)
        _lib.call("pic_dev_integrate_field", D.ptr(self.Es), D.ptr(out), self.Ng, self.dx, 1, D.stream())
        self.kernel_launches += 1
        return out.cpu().numpy()

    def check(self):
        D.check_range(self.range_err, "sheath step")
        if self.p2p is not None:
            self.p2p.check()

    def close(self):
        """Releases the peer-memory buffers (collective over the ranks); a no-op otherwise."""
        if self.p2p is not None:
            self.p2p.close()
            self.p2p = None
