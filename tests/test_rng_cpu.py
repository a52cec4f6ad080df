"""Host draw service (pypic_b200/rng.py + csrc/mt_host.cpp) against NumPy's legacy stream itself:
the MT19937 jump-ahead lands exactly where generating the skipped uniforms would, and the C
re-injection / thermostat draws are bit-identical to np.random.uniform / np.random.normal in the
reference's call order (PIC_L_DD.py:419-450).  No GPU involved."""
import numpy as np
import pytest

from pypic_b200.rng import LegacyDraws, sheath_step_draws


def _same_stream(a, b, n=64):
    sa, sb = a.get_state(), b.get_state()
    assert sa[3] == sb[3] and sa[4] == sb[4]                 # cached gaussian
    return np.array_equal(a.uniform(0, 1, n), b.uniform(0, 1, n)) and a.normal() == b.normal()


@pytest.mark.parametrize("n", [1, 311, 312, 313, 623, 624, 625, 9968, 19937, 250001, 3000017])
@pytest.mark.parametrize("warm", [0, 5, 1000])
def test_jump_equals_generation(n, warm):
    a, b = np.random.RandomState(42), np.random.RandomState(42)
    for r in (a, b):
        if warm:
            r.uniform(0, 1, warm); r.normal(size=3)          # odd number of normals: a cached gaussian is pending
    d = LegacyDraws(b)
    d.JUMP_MIN, d.CHUNK = 1, 256                             # force the jump path at every size
    a.uniform(0.0, 1.0, n)
    d.skip_uniforms(n)
    assert d.jumps == 1
    assert _same_stream(a, b)


def test_prefetched_jump_is_used_and_discarded_correctly():
    a, b = np.random.RandomState(7), np.random.RandomState(7)
    d = LegacyDraws(b)
    d.JUMP_MIN, d.CHUNK, d.MARGIN = 1000, 256, 100
    # hit: the next skip is at least as long as the prefetched one
    d.prefetch_skip(50000)
    a.uniform(0, 1, 49990); d.skip_uniforms(49990)
    assert d.prefetch_hits == 1 and _same_stream(a, b)
    # miss: the skip is shorter than the prefetched jump
    d.prefetch_skip(50000)
    a.uniform(0, 1, 30000); d.skip_uniforms(30000)
    assert d.prefetch_hits == 1 and _same_stream(a, b)
    # miss: somebody drew from the stream in between
    d.prefetch_skip(50000)
    a.normal(); b.normal()
    a.uniform(0, 1, 50000); d.skip_uniforms(50000)
    assert d.prefetch_hits == 1 and _same_stream(a, b)
    # short skips generate
    d.prefetch_skip(50000)
    a.uniform(0, 1, 10); d.skip_uniforms(10)
    assert _same_stream(a, b)


def test_reinjection_draws_match_numpy_call_order():
    a, b = np.random.RandomState(3), np.random.RandomState(3)
    a.normal(); b.normal()                                   # start with a cached gaussian
    L = 5e-4
    sig = np.where(np.arange(501) % 3 == 0, 1.3e6, 3.1e4)
    d = LegacyDraws(b)
    xd, ud, vd, wd = d.sheath_reinject(len(sig), sig, L)
    for k, s in enumerate(sig):
        assert xd[k] == a.uniform(0.0, L)
        assert ud[k] == a.normal(0.0, s) and vd[k] == a.normal(0.0, s) and wd[k] == a.normal(0.0, s)
    d.sheath_skip_foreign(77)
    for _ in range(77):
        a.uniform(0.0, 1.0); a.normal(0.0, 1.0); a.normal(0.0, 1.0); a.normal(0.0, 1.0)
    assert _same_stream(a, b)


@pytest.mark.parametrize("gamma", [0.0, 0.003, 0.2, 1.0])
def test_thermostat_draws_match_the_reference_loop(gamma):
    a, b = np.random.RandomState(5), np.random.RandomState(5)
    n_active, k_split, s0, s1 = 20011, 9000, 4.2e5, 9.8e3
    hits = []
    for k in range(n_active):                                # PIC_L_DD.py:420-426
        if a.uniform(0.0, 1.0) < gamma:
            s = s0 if k < k_split else s1
            hits.append((k, a.normal(0.0, s), a.normal(0.0, s), a.normal(0.0, s)))
    hk, hu, hv, hw = LegacyDraws(b).sheath_thermostat(n_active, k_split, gamma, s0, s1)
    assert len(hk) == len(hits)
    for (k, u, v, w), k2, u2, v2, w2 in zip(hits, hk, hu, hv, hw):
        assert (k, u, v, w) == (k2, u2, v2, w2)
    assert _same_stream(a, b)


def test_sharded_step_draws_consume_the_global_stream():
    """Two ranks with 3 and 2 dead slots, 1000 particles: each rank's draws are the reference's for
    its slots and both leave the stream where the single-process loop leaves it."""
    ref = np.random.RandomState(9)
    L, N = 1e-3, 1000
    sig = [1.0, 2.0, 3.0, 4.0, 5.0]
    ref.uniform(0, 1, N - 5)
    want = [(ref.uniform(0.0, L), ref.normal(0.0, s), ref.normal(0.0, s), ref.normal(0.0, s)) for s in sig]
    for rank, sl in ((0, slice(0, 3)), (1, slice(3, 5))):
        r = np.random.RandomState(9)
        xd, ud, vd, wd = sheath_step_draws(LegacyDraws(r), [3, 2], rank, N, np.array(sig[sl]), L)
        assert [tuple(t) for t in zip(xd, ud, vd, wd)] == want[sl]
        assert _same_stream(r, np.random.RandomState(9) if False else _advance(ref))


def _advance(r):
    c = np.random.RandomState(0)
    c.set_state(r.get_state())
    return c


def test_held_stream_gives_the_same_draws_and_is_written_back():
    """hold(): the state lives in the draw service during a time loop and returns to np.random at the end."""
    a, b = np.random.RandomState(21), np.random.RandomState(21)
    d = LegacyDraws(b)
    d.JUMP_MIN, d.CHUNK, d.MARGIN = 1000, 256, 100
    L = 2e-3
    with d.hold():
        for step in range(6):
            n_act, n_dead = 40000 - step, 3 + step
            a.uniform(0, 1, n_act)
            d.sheath_thermostat_skip(n_act)
            sig = np.full(n_dead, 7.0 + step)
            xd, ud, vd, wd = d.sheath_reinject(n_dead, sig, L)
            for k in range(n_dead):
                assert (xd[k], ud[k], vd[k], wd[k]) == (a.uniform(0.0, L), a.normal(0.0, sig[k]), a.normal(0.0, sig[k]),
                                                         a.normal(0.0, sig[k]))
            d.prefetch_skip(n_act)
        d.sheath_skip_foreign(2)
        for _ in range(2):
            a.uniform(0.0, 1.0); a.normal(); a.normal(); a.normal()
    assert d.prefetch_hits >= 4
    assert _same_stream(a, b)
