"""Host draw service: reproduces the reference's NumPy legacy MT19937 draw order
so that a seeded run re-injects exactly the particles the reference would.

The legacy stream is inherently sequential (polar legacy_gauss with a cached second
variate, data-dependent rejection), so for parity-sized runs the draws are made on
the host with ``np.random`` (global state by default -- exactly what the reference
uses) in the reference's call order and shipped to the device.  Benchmark-sized runs
use the device Philox generator instead (statistical parity only).
"""
import numpy as np


class LegacyDraws:
    def __init__(self, rng=None):
        self.rng = np.random if rng is None else rng

    # PIC_L_DD.py:419-450 -------------------------------------------------------------
    def sheath_thermostat_skip(self, n_active):
        """gamma == 0: the short-circuit ``and`` still draws one uniform per ACTIVE
        particle (PIC_L_DD.py:421); vectorised draws consume the same stream."""
        if n_active:
            self.rng.uniform(0.0, 1.0, int(n_active))

    def sheath_reinject(self, n_dead, sigma, L):
        """Per dead slot, in index order: x=uniform(0,L) then u,v,w=normal(0,sigma_i)
        (PIC_L_DD.py:433-436 / 443-446).  sigma: array of per-slot thermal speeds."""
        xd = np.empty(n_dead); ud = np.empty(n_dead); vd = np.empty(n_dead); wd = np.empty(n_dead)
        rng = self.rng
        for k in range(n_dead):
            s = sigma[k]
            xd[k] = rng.uniform(0.0, L)
            ud[k] = rng.normal(0.0, s)
            vd[k] = rng.normal(0.0, s)
            wd[k] = rng.normal(0.0, s)
        return xd, ud, vd, wd

    def sheath_skip_foreign(self, n_dead):
        """Advance the stream past the draws of dead slots owned by lower ranks."""
        rng = self.rng
        for _ in range(int(n_dead)):
            rng.uniform(0.0, 1.0)
            rng.normal(0.0, 1.0); rng.normal(0.0, 1.0); rng.normal(0.0, 1.0)


def sheath_step_draws(draws, counts, rank, N_global, sigma_local, L):
    """Sharded re-injection with stream parity: advances `draws` exactly as the reference's
    single process would for the GLOBAL particle list (PIC_L_DD.py:419-450: one thermostat
    uniform per active particle, then x,u,v,w per dead slot in index order) and returns only the
    draws of this rank's dead slots.  counts[r] = number of dead slots on rank r (rank order =
    index order because shards are contiguous index ranges); sigma_local = thermal speed of each
    of this rank's dead slots."""
    counts = [int(c) for c in counts]
    draws.sheath_thermostat_skip(int(N_global) - sum(counts))
    draws.sheath_skip_foreign(sum(counts[:rank]))
    out = draws.sheath_reinject(counts[rank], sigma_local, L)
    draws.sheath_skip_foreign(sum(counts[rank + 1:]))
    return out
