"""Spatial (slab) domain decomposition of the bounded two-species sheath (BASELINE config 5;
a capability the reference does not have -- SURVEY.md 8e).

Rank r owns the cells [c_r, c_{r+1}) of the global grid and the particles inside them, one
structure-of-arrays block per species.  The particle kernels are the ones of the
particle-decomposed path (pypic_b200/sheath.py: fused gather + CN push + walls + jh/j1 deposit
on the global node numbering), so what changes is the communication:

* **halo exchange + distributed field update** instead of an all-reduce of the whole grid: a rank's
  particles deposit only on its own nodes and on `guard` nodes either side, and every rank updates the
  field on that band only (csrc/slab_kernels.cu).  Per Picard iteration one all-gather carries each
  rank's message (the raw currents of the 2*guard+1 nodes around its two boundaries, the sum of all its
  raw deposits, four absorbed counts) and a second one the partial sums of the residual; neighbours
  compute identical bits on the nodes they share.  field="replicated" keeps round 1's scheme (guard strips
  + owned segments all-gathered, the whole grid updated on every rank);
* **particle migration**: every `sort_every` steps each species block is counting-sorted by cell
  (pic_dev_dd_sort_by_cell).  In a sorted block the particles that left the slab are contiguous
  runs at its two ends, ordered by destination rank, so the send buffers are views of the sorted
  array (no packing pass) and one all-to-all moves them; arrivals are placed in the headroom
  directly before / after the stayers (no copy of the bulk).  `guard` must cover the drift of
  `sort_every` steps;
* **re-injection** (PIC_L_DD.py:429-450) draws x uniformly over the WHOLE domain, so a revived
  particle usually belongs to another rank: the draws of a step are keyed by the global ordinal of
  the dead particle within its species (Philox; the particle set is therefore identical for any
  number of ranks) and REPLICATED -- after one all-gather of the dead counts every rank generates all
  draws of the step and keeps the ones that land in its slab; arrivals fill the rank's dead slots,
  left-over slots are closed by swap-removal.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, device as D
from .dist import Comm

epsilon0 = 8.854E-12
e = 1.602E-19
mp = 1.67E-27
me = 9.11E-31


def slab_bounds(Ng, world):
    """First cell of every rank's slab (world+1 entries; cells = Ng-1)."""
    cells = int(Ng) - 1
    return [r * cells // int(world) for r in range(int(world))] + [cells]


def exchange_plan(Ng, G, cb, rank):
    """Index plan of the per-iteration exchange of the accumulator [jh(Ng) | j1(Ng) | 4 counts | zero].

    Every rank contributes one message [left guard strip | right guard strip | owned segment (jh,
    j1) | 4 absorbed counts]; the strips of a rank are ITS deposits on its neighbours' nodes (G nodes
    below its first cell, G+1 nodes from its last cell boundary on).  One all-gather moves all
    messages; unpacking places the owned segments (`unpack`: where node i of jh / j1 sits in the
    gathered buffer) and adds every strip at its global position (`add_src` -> `add_dst`).
    Pure index arithmetic (NumPy), tested on the CPU by emulating the ranks."""
    W = len(cb) - 1
    ar = lambda a, b: np.arange(a, b, dtype=np.int64)
    both = lambda a, b: np.concatenate([ar(a, b), Ng + ar(a, b)])          # the same nodes of jh and of j1
    seg = [(cb[r], cb[r + 1] + (1 if r == W - 1 else 0)) for r in range(W)]   # the last rank also owns node Ng-1
    sl = max(b - a for a, b in seg)
    nL, nR = 2 * G, 2 * (G + 1)
    M = nL + nR + 2 * sl + 4                                          # message length
    PAD = 2 * Ng + 4                                                  # index of the always-zero slot of acc
    pack = np.full(M, PAD, dtype=np.int64)
    c0, c1 = cb[rank], cb[rank + 1]
    if rank > 0:
        pack[:nL] = both(c0 - G, c0)
    if rank < W - 1:
        pack[nL:nL + nR] = both(c1, c1 + G + 1)
    a, b = seg[rank]
    o = nL + nR
    pack[o:o + b - a] = ar(a, b); pack[o + sl:o + sl + b - a] = Ng + ar(a, b); pack[o + 2 * sl:] = 2 * Ng + ar(0, 4)
    src = np.zeros(2 * Ng, dtype=np.int64)
    add_dst, add_src = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)]
    for rr, (a2, b2) in enumerate(seg):
        base = rr * M
        src[a2:b2] = base + o + ar(0, b2 - a2); src[Ng + a2:Ng + b2] = base + o + sl + ar(0, b2 - a2)
        c0r, c1r = cb[rr], cb[rr + 1]
        if rr > 0:
            add_dst.append(both(c0r - G, c0r)); add_src.append(base + ar(0, nL))
        if rr < W - 1:
            add_dst.append(both(c1r, c1r + G + 1)); add_src.append(base + nL + ar(0, nR))
    return dict(M=M, pack=pack, unpack=src, add_dst=np.concatenate(add_dst), add_src=np.concatenate(add_src),
                counts=np.concatenate([rr * M + o + 2 * sl + ar(0, 4) for rr in range(W)]))


def swap_remove_plan(n, holes, n_arrivals):
    """Index plan that closes `holes` (ascending slot indices < n) of a block of n slots and adds
    n_arrivals new particles, moving as little as possible: arrivals fill holes first; left-over
    holes are filled from the block's tail (slots that are not holes themselves); left-over
    arrivals are appended.  Returns (arrival_dst, move_src, move_dst, new_n)."""
    holes = np.asarray(holes, dtype=np.int64)
    k = min(len(holes), int(n_arrivals))
    arrival_dst = list(holes[:k])
    rest = holes[k:]
    if len(rest) == 0:
        extra = int(n_arrivals) - k
        arrival_dst += list(range(n, n + extra))
        return np.asarray(arrival_dst, dtype=np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64), n + extra
    new_n = n - len(rest)
    hole_set = set(int(h) for h in rest)
    src = [s for s in range(new_n, n) if s not in hole_set]
    dst = [int(h) for h in rest if h < new_n]
    assert len(src) == len(dst)
    return (np.asarray(arrival_dst, dtype=np.int64), np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64), new_n)


class _Block:
    """One species on one rank: current state (x0,u0) as views into one of two allocations,
    the other allocation is the Picard / sort scratch."""

    def __init__(self, cap, dev, species):
        self.cap, self.dev, self.species = int(cap), dev, species
        self.X = [D.f64(self.cap, dev, True), D.f64(self.cap, dev, True)]
        self.U = [D.f64(self.cap, dev, True), D.f64(self.cap, dev, True)]
        self.cur, self.off, self.n = 0, 0, 0
        self.active = torch.ones(self.cap, dtype=torch.int8, device=dev)
        self.dead_idx = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.block_counts = torch.zeros(2 * (self.cap // 2048 + 2), dtype=torch.int64, device=dev)
        # absorption log of the particle kernels (pic_dev_dd_picard_iter4): int32 [count,0,0,0 | {slot,orig,iteration,0} x cap]
        self.dead_cap = int(max(1 << 16, self.cap // 64))
        self.dead_buf = torch.zeros(4 + 4 * self.dead_cap, dtype=torch.int32, device=dev)
        self.log_valid = False           # the log names every dead slot only after a step that started all-active

    # current / scratch views (the scratch always starts at slot 0 of the other allocation)
    @property
    def x0(self): return self.X[self.cur][self.off:]
    @property
    def u0(self): return self.U[self.cur][self.off:]
    @property
    def x1(self): return self.X[1 - self.cur]
    @property
    def u1(self): return self.U[1 - self.cur]

    def commit(self):
        self.cur, self.off = 1 - self.cur, 0


class SlabSheathSim:
    def __init__(self, N, Ng, dx, dt, p2c, q=(-e, e), m=(me, mp), kBT=(None, None), tol=1e-5, maxiter=20, seed=1,
                 comm=None, device=None, sort_every=8, guard=16, capacity=1.3, headroom=None, field="distributed"):
        self.dev = D.require_cuda(device)
        # field="distributed" (default): every rank updates E on its own nodes + guard nodes; per Picard iteration two
        # small all-gathers (boundary bands + partial sums, then the residual partials; csrc/slab_kernels.cu).
        # field="replicated": the round-1 scheme -- all-gather of the owned segments, every rank updates the whole
        # grid (kept for A/B runs and as a cross-check)
        if field not in ("distributed", "replicated"):
            raise ValueError("field must be 'distributed' or 'replicated'")
        self.field = field
        self.comm = comm if comm is not None else Comm()
        self.rank, self.world = self.comm.rank, self.comm.world
        self.N_global, self.Ng, self.dx, self.dt, self.p2c = int(N), int(Ng), float(dx), float(dt), float(p2c)
        self.L = dx * (Ng - 1)
        self.q, self.m, self.kBT = tuple(q), tuple(m), tuple(kBT)
        self.tol, self.maxiter, self.seed = float(tol), int(maxiter), int(seed)
        self.sort_every, self.G = int(sort_every), int(guard)
        self.cb = slab_bounds(Ng, self.world)
        self.c0, self.c1 = self.cb[self.rank], self.cb[self.rank + 1]
        if self.world > 1 and min(b - a for a, b in zip(self.cb, self.cb[1:])) <= 2 * self.G + 2:
            raise ValueError("slabs must be wider than 2*guard+2 cells")
        # Between two sorts a particle may deposit at most `guard` cells outside its slab: beyond that the halo
        # exchange does not carry its contribution.  Refuse configurations whose thermal drift (6 sigma of the
        # fastest species over sort_every steps) already exceeds the guard, and detect any leak at run time
        # (exchange_acc sums what lies outside the band; check() raises).
        if self.world > 1 and self.sort_every and all(k is not None for k in self.kBT):
            vmax = 6.0 * max(float(np.sqrt(self.kBT[s] / self.m[s])) for s in range(2))
            drift = vmax * self.dt * self.sort_every / self.dx
            if drift > self.G:
                raise ValueError("guard=%d cells cannot hold the drift of %d steps between sorts (6 sigma = %.1f cells): raise "
                                 "guard or lower sort_every" % (self.G, self.sort_every, drift))
        self.guard_leak = D.f64(1, self.dev, True)
        per = self.N_global // 2 // self.world + 1
        self.H = int(headroom) if headroom is not None else max(4096, per // 50)
        self.H += self.H % 2
        cap = int(per * capacity) + 2 * self.H + 1024
        self.blocks = [_Block(cap, self.dev, 0), _Block(cap, self.dev, 1)]
        g, dev = self.Ng, self.dev
        self.E0 = D.f64(g, dev, True); self.Es = D.f64(g, dev, True); self.E1 = D.f64(g, dev, True)
        self.j0 = D.f64(g, dev, True)
        self.acc = D.f64(2 * g + 5, dev, True)       # [jh | j1 | 4 absorbed counts | one always-zero slot]
        self.wall_cum = D.f64(4, dev, True)
        # [r, mean j1, EE, iterations | 4 words of reduction scratch | residual of every iteration of the step]
        self.stats = D.f64(8 + self.maxiter, dev, True)
        # enqueue-ahead Picard loop (see SheathSim.picard): the iterations the previous step needed are queued
        # without a host round trip each, behind a device flag the field kernel raises when the loop ends
        self.ctl = torch.zeros(1, dtype=torch.int32, device=dev)
        self.enqueue_ahead = True
        self._prev_k = 0
        self._absorbed_local = torch.zeros(4, dtype=torch.float64, device=dev)
        self.range_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sort_counts = torch.zeros(D.sort_counts_size(g), dtype=torch.int32, device=dev)
        self.cb_dev = torch.as_tensor(np.asarray(self.cb[1:-1], dtype=np.int64), device=dev)
        self.t = 0
        self.kernel_launches = 0
        self.stat = dict(migrated=0, exported=0, imported=0)
        self._plan = None
        self.local_dead = [0, 0]     # absorbed particles of each species on this rank in the last step
        self.profile = None          # set to a dict to accumulate wall-clock seconds per section of step()
        self.iter_events = None      # set to a list to record a CUDA-event pair per Picard iteration (both blocks)
        self.rprof = None            # set to a dict: synchronising wall-clock sections of reinject() (diagnostics)
        self.phase_events = None     # set to a list: CUDA events at the phase boundaries of every iteration (distributed field)
        # segment bookkeeping of the all-gather (owned nodes; the last rank also owns node Ng-1)
        self.seg = [(self.cb[r], self.cb[r + 1] + (1 if r == self.world - 1 else 0)) for r in range(self.world)]
        self.seglen = max(b - a for a, b in self.seg)
        # distributed field update: message / gather buffers (layout in csrc/slab_kernels.cu)
        lib = _lib.load()
        self.M = int(lib.pic_slab_message_len(self.G))
        self.msg = D.f64(self.M, dev, True); self.gath = D.f64(self.world * self.M, dev, True)
        self.part = D.f64(2, dev, True); self.gath2 = D.f64(2 * self.world, dev, True)
        self.work = D.f64(int(lib.pic_slab_work_len()), dev, True)
        self.b0, self.b1 = max(self.c0 - self.G, 0), min(self.c1 + self.G + 1, self.Ng)     # the band this rank computes
        self._leak_idx = None
        if self.world > 1 and getattr(self.comm, "enabled", False) and dist.is_initialized():
            self._warm_collectives()          # (emulated ranks -- tests -- have no process group)

    def _warm_collectives(self):
        """One call of every collective the step uses, with non-empty messages: NCCL sets up its point-to-point
        connections on first use (100-400 ms per kind, measured), which would otherwise land in whichever step first
        ships a re-injected particle or migrates one."""
        W, g = self.world, self.comm.group
        z = torch.zeros(2 * W, dtype=torch.float64, device=self.dev); o = torch.empty_like(z)
        dist.all_to_all_single(o, z, [2] * W, [2] * W, group=g)
        zi = torch.zeros(2 * W, dtype=torch.int64, device=self.dev); oi = torch.empty_like(zi)
        dist.all_to_all_single(oi, zi, group=g)
        dist.all_to_all([torch.empty(1, dtype=torch.float64, device=self.dev) for _ in range(W)],
                        [torch.zeros(1, dtype=torch.float64, device=self.dev) for _ in range(W)], group=g)
        dist.all_gather_into_tensor(oi, zi[:2], group=g)
        dist.all_gather_into_tensor(self.gath, self.msg, group=g)
        dist.all_gather_into_tensor(self.gath2, self.part, group=g)
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ helpers
    def _params(self, blk, n=None, sort=False):
        n = blk.n if n is None else n
        ns = n if (blk.species == 0 or sort) else 0          # sort keys: one species per block
        return _lib.DDParams(n, ns, self.Ng, 0, self.dx, self.dt, self.L, self.p2c, (C.c_double * 2)(*self.q),
                             (C.c_double * 2)(*self.m))

    def _sigma(self, sp):
        return float(np.sqrt(self.kBT[sp] / self.m[sp]))

    def _dest(self, x):
        """Owner rank of positions x (device tensor)."""
        cell = torch.clamp(torch.floor(x / self.dx).to(torch.int64), 0, self.Ng - 2)
        return torch.searchsorted(self.cb_dev, cell, right=True)

    def _all_to_all(self, send_list):
        """Variable-size all-to-all of fp64 vectors: exchanges the sizes, then the payloads."""
        W = self.world
        sizes = torch.tensor([t.numel() for t in send_list], dtype=torch.int64, device=self.dev)
        rsizes = torch.empty_like(sizes)
        dist.all_to_all_single(rsizes, sizes, group=self.comm.group)
        rs = [int(v) for v in rsizes.cpu().numpy()]
        recv = [torch.empty(n, dtype=torch.float64, device=self.dev) for n in rs]
        dist.all_to_all(recv, [t.contiguous() for t in send_list], group=self.comm.group)
        return recv

    # ------------------------------------------------------------------ state I/O
    def upload(self, x0, u0, E0=None):
        """GLOBAL host arrays (first half electrons, second half ions): every rank keeps the
        particles inside its slab."""
        h = self.N_global // 2
        lo, hi = self.c0 * self.dx, self.c1 * self.dx
        for sp, sl in ((0, slice(0, h)), (1, slice(h, self.N_global))):
            xs, us = np.asarray(x0[sl]), np.asarray(u0[sl])
            cell = np.clip(np.floor(xs / self.dx).astype(np.int64), 0, self.Ng - 2)
            keep = (cell >= self.c0) & (cell < self.c1)
            blk = self.blocks[sp]
            blk.n = int(keep.sum())
            assert blk.n + 2 * self.H < blk.cap, "slab block capacity too small"
            blk.cur, blk.off = 0, 0
            blk.X[0][:blk.n].copy_(torch.as_tensor(np.ascontiguousarray(xs[keep])))
            blk.U[0][:blk.n].copy_(torch.as_tensor(np.ascontiguousarray(us[keep])))
            blk.active.fill_(1); blk.log_valid = False
        if E0 is not None:
            self.E0.copy_(torch.as_tensor(np.ascontiguousarray(E0)))

    def init_device(self, seed=None):
        """Uniform density, Maxwellian velocities, generated in place (Philox keyed by rank)."""
        seed = self.seed if seed is None else seed
        h = self.N_global // 2
        lo, hi = self.c0 * self.dx, self.c1 * self.dx
        for sp, blk in enumerate(self.blocks):
            blk.n = h * (self.c1 - self.c0) // (self.Ng - 1)
            blk.cur, blk.off = 0, 0
            sig = (self._sigma(sp), self._sigma(sp))
            _lib.call("pic_dev_init_uniform_maxwellian", D.ptr(blk.X[0]), D.ptr(blk.U[0]), None, None, blk.n, blk.n,
                      max(lo, 1e-12 * self.L), min(hi, self.L * (1 - 1e-12)), C.byref((C.c_double * 2)(*sig)),
                      C.byref((C.c_double * 2)(0., 0.)), seed, 100 + sp, self.rank * (1 << 40), D.stream())
            blk.active.fill_(1); blk.log_valid = False

    def gather_particles(self):
        """All particles of both species on every rank (tests): lists of numpy arrays [x_e, u_e, x_i, u_i]."""
        out = []
        for blk in self.blocks:
            x = blk.x0[:blk.n].cpu().numpy(); u = blk.u0[:blk.n].cpu().numpy(); a = blk.active[:blk.n].cpu().numpy()
            if self.world > 1:
                parts = [None] * self.world
                dist.all_gather_object(parts, (x, u, a), group=self.comm.group)
                x = np.concatenate([p[0] for p in parts]); u = np.concatenate([p[1] for p in parts])
                a = np.concatenate([p[2] for p in parts])
            out += [x, u, a]
        return out

    # ------------------------------------------------------------------ re-injection
    def reinject(self):
        """PIC_L_DD.py:429-450 on slabs: every dead slot gets (x ~ U(0,L), u ~ N(0, sigma_s)) drawn for the
        GLOBAL ordinal of that dead particle within its species (Philox keyed by the ordinal, so the particle set
        does not depend on the number of ranks).  The draws are REPLICATED: after one all-gather of the per-rank
        dead counts every rank generates all draws of the step (a few thousand) and keeps the ones that land in
        its slab -- no size exchange, no payload exchange.  Arrivals fill the rank's dead slots first; left-over
        slots are closed by swap-removal, left-over arrivals appended."""
        st = D.stream()
        W, me_ = self.world, self.rank
        dead = []
        import time
        t_last = [time.perf_counter()]

        def tick(name):
            if self.rprof is not None:
                torch.cuda.synchronize(); t = time.perf_counter()
                self.rprof[name] = self.rprof.get(name, 0.0) + t - t_last[0]; t_last[0] = t
        for sp, blk in enumerate(self.blocks):
            # the kernels count what they absorb (saved per rank by picard()): a species that lost nothing here
            # needs no look at its flags (always the case on interior ranks)
            if self.local_dead[sp] == 0 and self.t > 0:
                dead.append(0)
                continue
            # the slots the particle kernels logged while absorbing (a few hundred), sorted into index order;
            # the flag scan only runs without a valid log (first step, overflow)
            nd = -1
            if blk.log_valid:
                nd = int(D.read_raw(blk.dead_buf, 1, np.int32)[0])
                if nd > blk.dead_cap:
                    nd = -1
                elif nd:
                    blk.dead_idx[:nd] = torch.sort(blk.dead_buf[4:4 + 4 * nd].view(nd, 4)[:, 0]).values
            if nd < 0:
                _lib.call("pic_dev_compact_flags", D.ptr(blk.active), blk.n, 0, D.ptr(blk.dead_idx), D.ptr(blk.count),
                          D.ptr(blk.block_counts), st)
                nd = int(D.read_raw(blk.count, 1, np.int64)[0])
                self.kernel_launches += 3
            dead.append(nd)
        tick("dead_slots")
        if W > 1:
            t = torch.tensor(dead, dtype=torch.int64, device=self.dev)
            allc = torch.empty(W * 2, dtype=torch.int64, device=self.dev)
            dist.all_gather_into_tensor(allc, t, group=self.comm.group)
            allc = allc.cpu().numpy().reshape(W, 2)                       # [rank, species]
        else:
            allc = np.asarray([dead])
        tick("count_exchange")
        for sp, blk in enumerate(self.blocks):
            total = int(allc[:, sp].sum())
            if total == 0:
                continue
            nd = dead[sp]
            first = int(allc[:me_, sp].sum())                # global ordinal of this rank's first dead particle
            xd = D.f64(total, self.dev); ud = D.f64(total, self.dev)
            sig = (self._sigma(sp), self._sigma(sp))
            _lib.call("pic_dev_init_uniform_maxwellian", D.ptr(xd), D.ptr(ud), None, None, total, total, 0.0, self.L,
                      C.byref((C.c_double * 2)(*sig)), C.byref((C.c_double * 2)(0., 0.)), self.seed,
                      1000 + 2 * self.t + sp, 0, st)
            self.kernel_launches += 1
            if W == 1:
                idx = blk.dead_idx[:nd].to(torch.int64)
                blk.x0[idx] = xd; blk.u0[idx] = ud; blk.active[idx] = 1
                continue
            mine = self._dest(xd) == me_
            # ONE device->host read per species: [arrivals, own draws that stay | the rank's dead slots]
            head = torch.stack([mine.sum(), mine[first:first + nd].sum()]).to(torch.int64)
            info = torch.cat([head, blk.dead_idx[:nd].to(torch.int64)]).cpu().numpy()
            n_arr, n_stay = int(info[0]), int(info[1])
            holes = info[2:]
            ax, au = xd[mine], ud[mine]
            adst, msrc, mdst, new_n = swap_remove_plan(blk.n, holes, n_arr)
            assert blk.off + new_n + 16 < blk.cap, "slab block capacity exhausted"
            if len(msrc):
                ms = torch.as_tensor(msrc, device=self.dev); md = torch.as_tensor(mdst, device=self.dev)
                blk.x0[md] = blk.x0[ms]; blk.u0[md] = blk.u0[ms]; blk.active[md] = blk.active[ms]
            if len(adst):
                ad = torch.as_tensor(adst, device=self.dev)
                blk.x0[ad] = ax; blk.u0[ad] = au; blk.active[ad] = 1
            blk.n = new_n
            self.stat["exported"] += nd - n_stay; self.stat["imported"] += n_arr - n_stay
        tick("draw_and_place")

    # ------------------------------------------------------------------ sort + migration
    def migrate_sort(self):
        st = D.stream()
        W, me_ = self.world, self.rank
        for blk in self.blocks:
            n, H = blk.n, self.H
            xs, us = blk.x1[H:], blk.u1[H:]
            P = self._params(blk, n, sort=True)
            _lib.call("pic_dev_dd_sort_by_cell", C.byref(P), D.ptr(blk.x0), D.ptr(blk.u0), None, None, D.ptr(xs), D.ptr(us),
                      None, None, D.ptr(self.sort_counts), st)
            self.kernel_launches += 3
            # after the scatter cursor[key] = end of key's run = start of the next key's run
            if W > 1:
                ends = self.sort_counts[torch.as_tensor([c - 1 for c in self.cb[1:-1]], device=self.dev)].cpu().numpy()
                offs = np.concatenate([[0], ends.astype(np.int64), [n]])
            else:
                offs = np.asarray([0, n], dtype=np.int64)
            start, stay = H + int(offs[me_]), int(offs[me_ + 1] - offs[me_])
            if W > 1:
                send_x = [xs[offs[r]:offs[r + 1]] if r != me_ else xs[:0] for r in range(W)]
                send_u = [us[offs[r]:offs[r + 1]] if r != me_ else us[:0] for r in range(W)]
                rx, ru = self._all_to_all(send_x), self._all_to_all(send_u)
                lo_x, hi_x = torch.cat(rx[:me_] + [xs[:0]]), torch.cat(rx[me_ + 1:] + [xs[:0]])
                lo_u, hi_u = torch.cat(ru[:me_] + [us[:0]]), torch.cat(ru[me_ + 1:] + [us[:0]])
                nlo, nhi = lo_x.numel(), hi_x.numel()
                assert nlo <= start, "headroom too small for the arrivals from lower ranks"
                Xs, Us = blk.x1, blk.u1
                if nlo:
                    Xs[start - nlo:start] = lo_x; Us[start - nlo:start] = lo_u
                if nhi:
                    Xs[start + stay:start + stay + nhi] = hi_x; Us[start + stay:start + stay + nhi] = hi_u
                self.stat["migrated"] += int(n - stay)
                start, n = start - nlo, nlo + stay + nhi
            if start % 2:                       # keep the block 16-byte aligned for the TMA / 128-bit paths
                Xs, Us = blk.x1, blk.u1
                Xs[start - 1] = Xs[start + n - 1]; Us[start - 1] = Us[start + n - 1]
                start -= 1
            assert start + n + 16 < blk.cap, "slab block capacity exhausted"
            blk.cur, blk.off, blk.n = 1 - blk.cur, start, n
            blk.active[:n] = 1

    # ------------------------------------------------------------------ halo exchange
    def _build_exchange_plan(self):
        """Device copies of the index plan of exchange_plan(); allocates the message buffers."""
        pl = exchange_plan(self.Ng, self.G, self.cb, self.rank)
        t = lambda v: torch.as_tensor(v, device=self.dev)
        self.sendbuf = D.f64(pl["M"], self.dev, True)
        self.gathbuf = D.f64(self.world * pl["M"], self.dev, True)
        lo, hi = max(self.c0 - self.G, 0), min(self.c1 + self.G + 1, self.Ng)
        out = np.concatenate([np.arange(0, lo), np.arange(hi, self.Ng)]).astype(np.int64)
        pl["outside"] = np.concatenate([out, self.Ng + out])
        return {k: t(v) for k, v in pl.items() if k != "M"}

    def exchange_acc(self):
        """field="replicated": halo exchange of the guard strips + completion of the grid, one collective per
        Picard iteration (see _build_exchange_plan).  self.acc has one extra, always-zero slot."""
        W = self.world
        if W == 1:
            return
        if self._plan is None:
            self._plan = self._build_exchange_plan()
        pl, Ng = self._plan, self.Ng
        acc = self.acc
        # anything this rank deposited OUTSIDE its slab + guard band would be dropped by the unpack below
        self.guard_leak += acc[pl["outside"]].abs().sum()
        torch.index_select(acc, 0, pl["pack"], out=self.sendbuf)
        dist.all_gather_into_tensor(self.gathbuf, self.sendbuf, group=self.comm.group)
        torch.index_select(self.gathbuf, 0, pl["unpack"], out=acc[:2 * Ng])
        acc.index_add_(0, pl["add_dst"], self.gathbuf[pl["add_src"]])
        acc[2 * Ng:2 * Ng + 4] = self.gathbuf[pl["counts"]].view(W, 4).sum(0)

    def _all_gather(self, out, src):
        if self.world == 1:
            out.copy_(src)
        else:
            dist.all_gather_into_tensor(out, src, group=self.comm.group)

    def _slab_args(self):
        return (C.byref(self._params(self.blocks[0])), self.c0, self.c1, self.G, self.rank, self.world)

    # The three device phases of one Picard iteration with field="distributed" (the collectives between them are
    # picard()'s; tests drive several emulated ranks on one GPU through these phases)
    def iter_particles(self, j):
        """Particle kernels of both species blocks, then the boundary bands / partial sums -> self.msg."""
        st = D.stream()
        ev = None
        if self.iter_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        for blk in self.blocks:
            if blk.n:
                _lib.call("pic_dev_dd_picard_iter4", C.byref(self._params(blk)), D.ptr(blk.x0), D.ptr(blk.u0), D.ptr(blk.x1),
                          D.ptr(blk.x1), D.ptr(blk.u1), D.ptr(blk.active), D.ptr(self.Es), D.ptr(self.acc), 1 if j == 0 else 0,
                          D.ptr(self.range_err), D.ptr(self.ctl), D.ptr(blk.dead_buf), blk.dead_cap, None, j, st)
                self.kernel_launches += 1
        if ev is not None:
            ev[1].record()
            self.iter_events.append(ev)
        Ng = self.Ng
        if self.field == "distributed":
            # (also adds this rank's own absorptions of the iteration to _absorbed_local)
            _lib.call("pic_dev_slab_pack", *self._slab_args(), D.ptr(self.acc), D.ptr(self.msg), D.ptr(self.work),
                      D.ptr(self._absorbed_local), D.ptr(self.ctl), st)
            self.kernel_launches += 1
        else:
            self._absorbed_local += self.acc[2 * Ng:2 * Ng + 4]   # this rank's own absorptions (before the exchange)

    def iter_field(self):
        """self.gath (all ranks' messages) -> E1, Es, j0 on the band; residual / energy partials -> self.part."""
        _lib.call("pic_dev_slab_field_update", *self._slab_args(), D.ptr(self.acc), D.ptr(self.gath), D.ptr(self.wall_cum),
                  D.ptr(self.E0), D.ptr(self.Es), D.ptr(self.E1), D.ptr(self.j0), D.ptr(self.part), D.ptr(self.work),
                  D.ptr(self.ctl), D.stream())
        self.kernel_launches += 1

    def iter_finish(self):
        """self.gath2 (all ranks' partials) -> counts, statistics, loop flag."""
        _lib.call("pic_dev_slab_finish", *self._slab_args(), D.ptr(self.gath), D.ptr(self.gath2), D.ptr(self.wall_cum),
                  D.ptr(self.stats), D.ptr(self.stats) + 8 * 8, D.ptr(self.ctl), self.tol, self.maxiter, D.stream())
        self.kernel_launches += 1

    def begin_step(self):
        if self.stats.numel() < 8 + self.maxiter:
            self.stats = D.f64(8 + self.maxiter, self.dev, True)
        self.Es.copy_(self.E0)
        self.wall_cum.zero_(); self.stats.zero_(); self.ctl.zero_()
        self._absorbed_local.zero_()

    def outcome(self):
        s_ = D.read_f64(self.stats, 8 + self.maxiter)
        k_ = int(s_[3])
        return k_, [float(v) for v in s_[8:8 + k_]]

    def end_step(self, k):
        self._prev_k = k
        if k > 0:
            for blk in self.blocks:
                blk.commit()
            self.E0, self.E1 = self.E1, self.E0
        a = self._absorbed_local.cpu().numpy()                      # [left e, left i, right e, right i]
        self.local_dead = [int(round(a[0] + a[2])), int(round(a[1] + a[3]))]
        if self.field == "distributed" and self.world > 1:
            # deposits outside the band are never exchanged (nor cleared): one look per step
            if self._leak_idx is None:
                out = np.concatenate([np.arange(0, self.b0), np.arange(self.b1, self.Ng)]).astype(np.int64)
                self._leak_idx = torch.as_tensor(np.concatenate([out, self.Ng + out]), device=self.dev)
            self.guard_leak += self.acc[self._leak_idx].abs().sum()

    def gather_field(self, which="E0"):
        """The complete grid array (E0 or j0) on every rank, assembled from the owned segments (diagnostics / tests;
        with field="replicated" every rank already holds it)."""
        a = getattr(self, which)
        if self.field == "replicated" or self.world == 1:
            return a.clone()
        buf = D.f64(self.seglen, self.dev, True)
        o0, o1 = self.seg[self.rank]
        buf[:o1 - o0] = a[o0:o1]
        allb = D.f64(self.world * self.seglen, self.dev, True)
        dist.all_gather_into_tensor(allb, buf, group=self.comm.group)
        out = D.f64(self.Ng, self.dev, True)
        for r, (s0, s1) in enumerate(self.seg):
            out[s0:s1] = allb[r * self.seglen:r * self.seglen + s1 - s0]
        return out

    # ------------------------------------------------------------------ one timestep
    def picard(self):
        """PIC_L_DD.py:452-545 on slabs.  Every launch of an iteration (particle kernels, field kernels) is
        guarded by the device flag `ctl`; the exchanges between them are unguarded library calls, which move
        stale messages once the loop has ended (nobody reads them), so queued iterations behind the end of the
        loop are harmless no-ops."""
        st = D.stream()
        self.begin_step()
        Pg = self._params(self.blocks[0])
        rhist = D.ptr(self.stats) + 8 * 8

        def mark():
            if self.phase_events is not None:      # diagnostics: 6 marks per iteration (tools/slab_phases.py)
                e = torch.cuda.Event(enable_timing=True); e.record(); self.phase_events.append(e)

        def launch(j):
            mark()
            self.iter_particles(j)
            if self.field == "distributed":
                mark(); self._all_gather(self.gath, self.msg)
                mark(); self.iter_field()
                mark(); self._all_gather(self.gath2, self.part)
                mark(); self.iter_finish()
                mark()
                return
            if self.profile is not None:
                import time
                torch.cuda.synchronize(); t0 = time.perf_counter()
                self.exchange_acc()
                torch.cuda.synchronize(); self.profile["exchange"] = self.profile.get("exchange", 0.0) + time.perf_counter() - t0
            else:
                self.exchange_acc()
            _lib.call("pic_dev_dd_field_update2", C.byref(Pg), D.ptr(self.acc), D.ptr(self.wall_cum), D.ptr(self.E0),
                      D.ptr(self.Es), D.ptr(self.E1), D.ptr(self.j0), D.ptr(self.stats), None, rhist, D.ptr(self.ctl),
                      self.tol, self.maxiter, st)
            self.kernel_launches += 1

        outcome = self.outcome
        k, hist, queued = 0, [], 0
        if self.enqueue_ahead and self._prev_k and self.profile is None:
            for j in range(min(self._prev_k, self.maxiter)):
                launch(j); queued += 1
            k, hist = outcome()
        r = hist[-1] if hist else 1.0
        while (r > self.tol) and (k < self.maxiter) and k == queued:
            launch(k); queued += 1
            k, hist = outcome()
            r = hist[-1]
        self.end_step(k)
        return k, r

    def step(self):
        if self.profile is not None:
            return self._step_profiled()
        self.reinject()
        self._reset_logs()
        if self.sort_every and self.t % self.sort_every == 0:
            self.migrate_sort()
        out = self.picard()
        self.t += 1
        return out

    def _reset_logs(self):
        """Every slot is alive after the re-injection: the absorption logs of the coming step start empty."""
        for blk in self.blocks:
            blk.dead_buf[:4].zero_()
            blk.log_valid = True

    def _step_profiled(self):
        """step() with host-side wall-clock sections (synchronising; diagnostics only)."""
        import time
        pr = self.profile

        def section(name, fn):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize(); pr[name] = pr.get(name, 0.0) + time.perf_counter() - t0
            return out
        section("reinject", self.reinject)
        self._reset_logs()
        if self.sort_every and self.t % self.sort_every == 0:
            section("migrate_sort", self.migrate_sort)
        out = section("picard", self.picard)
        self.t += 1
        return out

    def check(self):
        D.check_range(self.range_err, "slab sheath step")
        leak = float(self.guard_leak.item())
        if leak != 0.0:
            self.guard_leak.zero_()
            raise _lib.PicError(_lib.PIC_ERR_RANGE, "slab decomposition: particles deposited beyond the %d guard cells of their "
                                "slab (the halo exchange dropped that current): raise guard or lower sort_every" % self.G)

    def local_particles(self):
        return sum(b.n for b in self.blocks)
