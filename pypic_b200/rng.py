"""Host draw service: reproduces the reference's NumPy legacy MT19937 draw order
so that a seeded run re-injects exactly the particles the reference would.

The reference (PIC_L_DD.py:419-450) consumes the global legacy stream in particle index
order: one uniform per ACTIVE particle for the thermostat -- even at gamma == 0, Python
still evaluates the right operand of ``active[i]==1 and np.random.uniform(0,1) < gamma`` --
then x = uniform(0,L) and u,v,w = normal(0,sigma) per dead slot.  The draws that matter
(a few hundred per step) are made by pypic_b200/csrc/mt_host.cpp bit-identically to
np.random; the thermostat uniforms of a large run are SKIPPED by an MT19937 jump-ahead
(polynomial t^J mod the characteristic polynomial applied to the state, ~1 ms whatever J)
instead of being generated, and the jump for the next step is started in a background
thread as soon as this step's draws are done, so it overlaps the GPU's Picard loop.
Benchmark-sized runs that do not need stream parity use the device Philox generator.
"""
import ctypes as C
import queue
import threading

import numpy as np

from . import _lib


class _Done(threading.Event):
    """What skip_uniforms waits on (the interface of Thread.join)."""

    def join(self):
        self.wait()


_jobs = None


def _worker():
    """One persistent daemon thread runs the prefetched jumps (starting a thread per step costs more
    than the jump); the C call releases the GIL, so the jump overlaps the host's kernel launches."""
    global _jobs
    if _jobs is None:
        _jobs = queue.SimpleQueue()

        def loop():
            while True:
                job = _jobs.get()
                try:
                    p = C.c_int32(job["pos"])
                    poly = job["owner"]._poly(2 * job["n"])          # a cache miss costs 0.5 ms: paid here, not by the time loop
                    _lib.call("pic_mt_jump", job["key1"].ctypes.data, C.byref(p), poly.ctypes.data)
                    job["pos1"] = p.value
                except Exception:                       # leaves pos1 = 0 with an unchanged key: detected below
                    job["n"] = -1
                finally:
                    job["thread"].set()
        threading.Thread(target=loop, daemon=True).start()
    return _jobs


class LegacyDraws:
    JUMP_MIN = 200000        # uniforms; below this plain generation is cheaper than the jump
    CHUNK = 4096             # jump polynomials are cached per multiple of CHUNK uniforms; the rest is generated
                             # (~3 ns per uniform; a polynomial that is not cached costs 0.5 ms)
    MARGIN = 1024            # the prefetched jump stops this many uniforms short of the expected skip

    def __init__(self, rng=None):
        self.rng = np.random if rng is None else rng
        self._polys = {}
        self._pref = None
        self._held = None        # the stream's state while hold() owns it
        self.jumps = 0           # statistics: jumps applied / prefetched jumps used
        self.prefetch_hits = 0

    # ------------------------------------------------------------------ legacy state <-> C
    def _get(self):
        if self._held is not None:
            return self._held
        s = self.rng.get_state()
        if s[0] != "MT19937":
            raise ValueError("the legacy draw service needs an MT19937 stream, got %r" % (s[0],))
        return [np.array(s[1], dtype=np.uint32), C.c_int32(int(s[2])), C.c_int32(int(s[3])), C.c_double(float(s[4]))]

    def _put(self, st):
        if self._held is not None:
            self._held = st
            return
        self.rng.set_state(("MT19937", st[0], int(st[1].value), int(st[2].value), float(st[3].value)))

    def _uniforms(self, n):
        """n plain np.random.uniform() draws, discarded."""
        if self._held is not None:
            st = self._held
            _lib.load().pic_mt_skip(st[0].ctypes.data, C.byref(st[1]), 2 * int(n))
        else:
            self.rng.uniform(0.0, 1.0, int(n))

    def hold(self):
        """Context manager: for the duration of a time loop the stream's state lives in this object
        (C arrays) instead of being fetched from / stored to np.random around every call -- the state is
        written back on exit.  Nothing else may draw from the stream meanwhile."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            if self._held is not None:
                yield self
                return
            self._held = self._get()
            try:
                yield self
            finally:
                st, self._held = self._held, None
                if self._pref is not None:
                    self._pref["thread"].join()
                    self._pref = None
                self._put(st)
        return cm()

    def _poly(self, nwords):
        g = self._polys.get(nwords)
        if g is None:
            g = np.zeros(624, dtype=np.uint32)
            _lib.call("pic_mt_jump_poly", C.c_uint64(nwords), g.ctypes.data)
            if len(self._polys) > 64:
                self._polys.clear()
            self._polys[nwords] = g
        return g

    def skip_uniforms(self, n):
        """Advance the stream past n np.random.uniform() draws (two 32-bit words each)."""
        n = int(n)
        if n <= 0:
            return
        pref, self._pref = self._pref, None
        if pref is not None:
            pref["thread"].join()
        if n < self.JUMP_MIN:
            self._uniforms(n)
            return
        st = self._get()
        # the prefetched jump started from the stream's state at that time: usable if the stream has not moved
        # (guaranteed while hold() owns it: nothing else can draw; otherwise compare the snapshot)
        if (pref is not None and 0 < pref["n"] <= n and pref["pos"] == st[1].value and
                (pref["held"] and self._held is not None or np.array_equal(pref["key0"], st[0]))):
            st = [pref["key1"], C.c_int32(pref["pos1"]), st[2], st[3]]
            done = pref["n"]
            self.prefetch_hits += 1
        else:
            done = (n // self.CHUNK) * self.CHUNK
            _lib.call("pic_mt_jump", st[0].ctypes.data, C.byref(st[1]), self._poly(2 * done).ctypes.data)
        self.jumps += 1
        self._put(st)
        if n > done:
            self._uniforms(n - done)

    def prefetch_skip(self, n_expected):
        """Start the jump for the NEXT skip_uniforms() now, from the current state, in a thread (the C
        call releases the GIL).  It stops MARGIN uniforms short of the expectation; skip_uniforms
        generates the remainder, or discards the result if the stream moved or the skip is shorter."""
        n = ((int(n_expected) - self.MARGIN) // self.CHUNK) * self.CHUNK
        if n < self.JUMP_MIN:
            return
        st = self._get()
        held = self._held is not None
        job = dict(n=n, key0=None if held else st[0].copy(), pos=st[1].value, key1=st[0].copy(), pos1=0, owner=self,
                   thread=_Done(), held=held)
        _worker().put(job)
        self._pref = job

    # PIC_L_DD.py:419-450 -------------------------------------------------------------
    def sheath_thermostat_skip(self, n_active):
        """gamma == 0: the short-circuit ``and`` still draws one uniform per ACTIVE particle
        (PIC_L_DD.py:421)."""
        self.skip_uniforms(n_active)

    def sheath_thermostat(self, n_active, k_split, gamma, sigma0, sigma1):
        """gamma != 0 (PIC_L_DD.py:419-427): uniforms in index order over the active particles, every
        u < gamma followed by three normals.  Returns (ordinals of the hits among the active
        particles, u, v, w draws); sigma0 applies to ordinals < k_split."""
        cap = max(1024, int(1.5 * gamma * n_active) + 1024)
        while True:
            st = self._get()
            hk = np.empty(cap, dtype=np.int64)
            hu, hv, hw = np.empty(cap), np.empty(cap), np.empty(cap)
            nh = C.c_int64(0)
            lib = _lib.load()
            rc = lib.pic_mt_sheath_thermostat(st[0].ctypes.data, C.byref(st[1]), C.byref(st[2]), C.byref(st[3]), int(n_active),
                                              int(k_split), float(gamma), float(sigma0), float(sigma1), cap, hk.ctypes.data,
                                              hu.ctypes.data, hv.ctypes.data, hw.ctypes.data, C.byref(nh))
            if rc == 0:
                self._put(st)
                n = int(nh.value)
                return hk[:n], hu[:n], hv[:n], hw[:n]
            if nh.value <= cap:
                raise _lib.PicError(rc, lib.pic_last_error().decode())
            cap = int(nh.value) + 1024            # the state was left untouched: retry with room for every hit

    def sheath_reinject(self, n_dead, sigma, L):
        """Per dead slot, in index order: x=uniform(0,L) then u,v,w=normal(0,sigma_i)
        (PIC_L_DD.py:433-436 / 443-446).  sigma: array of per-slot thermal speeds."""
        n_dead = int(n_dead)
        out = np.empty((4, n_dead))              # one allocation: rows x, u, v, w
        if n_dead:
            sg = sigma if (isinstance(sigma, np.ndarray) and sigma.dtype == np.float64 and sigma.shape == (n_dead,)
                           and sigma.flags.c_contiguous) else \
                np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float64), (n_dead,)))
            st = self._get()
            p0 = out.ctypes.data
            _lib.call("pic_mt_sheath_draws", st[0].ctypes.data, C.byref(st[1]), C.byref(st[2]), C.byref(st[3]), n_dead,
                      sg.ctypes.data, float(L), p0, p0 + 8 * n_dead, p0 + 16 * n_dead, p0 + 24 * n_dead)
            self._put(st)
        return out[0], out[1], out[2], out[3]

    def sheath_skip_foreign(self, n_dead):
        """Advance the stream past the draws of dead slots owned by other ranks."""
        if int(n_dead) > 0:
            st = self._get()
            _lib.call("pic_mt_sheath_draws", st[0].ctypes.data, C.byref(st[1]), C.byref(st[2]), C.byref(st[3]), int(n_dead),
                      None, 1.0, None, None, None, None)
            self._put(st)


def sheath_step_draws(draws, counts, rank, N_global, sigma_local, L):
    """Sharded re-injection with stream parity (gamma == 0): advances `draws` exactly as the
    reference's single process would for the GLOBAL particle list (PIC_L_DD.py:419-450: one
    thermostat uniform per active particle, then x,u,v,w per dead slot in index order) and returns
    only the draws of this rank's dead slots.  counts[r] = number of dead slots on rank r (rank order
    = index order because shards are contiguous index ranges); sigma_local = thermal speed of each of
    this rank's dead slots, in index order."""
    counts = [int(c) for c in counts]
    draws.sheath_thermostat_skip(int(N_global) - sum(counts))
    return sheath_reinject_draws(draws, counts, rank, sigma_local, L)


def sheath_reinject_draws(draws, counts, rank, sigma_local, L):
    """The re-injection part alone (after the thermostat consumed its share of the stream)."""
    counts = [int(c) for c in counts]
    draws.sheath_skip_foreign(sum(counts[:rank]))
    out = draws.sheath_reinject(counts[rank], sigma_local, L)
    draws.sheath_skip_foreign(sum(counts[rank + 1:]))
    return out
