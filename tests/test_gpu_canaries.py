"""Out-of-bounds WRITE check of our own (compute-sanitizer is not available on the GPU pool):
every device buffer of a sheath simulation is re-homed inside a larger allocation filled with a
sentinel, whole steps run (default and reproducible build: TMA-staged Picard kernel, tail kernel,
field kernel, Philox re-injection, both counting-sort paths, the stable radix sort, repair pass),
and the sentinels on both sides of every buffer must be untouched."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import np_oracle as O

PAD = 2048            # elements on each side; keeps the 16-byte alignment the TMA path needs
SENT = {"torch.float64": -12345.678, "torch.int8": 77, "torch.int32": 0x5a5a5a5a, "torch.int64": 0x5a5a5a5a5a5a5a5a}


def _rehome(sim, names):
    import torch
    homes = []
    for nm in names:
        t = getattr(sim, nm)
        if t is None:
            continue
        big = torch.full((t.numel() + 2 * PAD,), SENT[str(t.dtype)], dtype=t.dtype, device=t.device)
        big[PAD:PAD + t.numel()].copy_(t)
        setattr(sim, nm, big[PAD:PAD + t.numel()])
        homes.append((nm, big, t.numel()))
    return homes


@pytest.mark.parametrize("deposit", ["window", "window-det"])
@pytest.mark.parametrize("N,Ng", [(3 * 16384 + 777, 257), (16384 + 1, 51), (5000, 129)])
def test_whole_steps_leave_the_canaries_around_every_buffer_intact(deposit, N, Ng):
    import torch
    from pypic_b200.sheath import SheathSim
    dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1)
    kT = O.kb * 116000.
    s = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=False, deposit=deposit, rng="philox", seed=1,
                  sort_every=2)
    homes = _rehome(s, ["x0", "u0", "x1", "x1b", "u1", "active", "E0", "Es", "E1", "Es_prev", "j0", "acc", "wall_cum",
                        "stats", "range_err", "sort_counts", "sort_scratch", "ctl", "scalar"])
    gen = torch.Generator(device=s.dev); gen.manual_seed(5)
    s.x0.uniform_(0.0, 1.0, generator=gen).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
    s.u0.normal_(0.0, 1.0, generator=gen)
    s.u0[:s.n_split].mul_(float(np.sqrt(kT / O.me))); s.u0[s.n_split:].mul_(float(np.sqrt(kT / O.mp)))
    s.E0.normal_(0.0, 1e4, generator=gen)
    absorbed = 0
    for step in range(6):
        if step == 4:
            s._prev_hist = [1e30] * (len(s._prev_hist) + 2)      # forces light iterations, no-op launches and the repair pass
        s.step()
        absorbed += int((s.active != 1).sum())
        s.diagnostics()
    s.check()
    assert absorbed > 0 and s.u_repairs >= 1
    torch.cuda.synchronize()
    for nm, big, n in homes:
        sent = SENT[str(big.dtype)]
        lo, hi = big[:PAD], big[PAD + n:]
        assert bool((lo == sent).all()) and bool((hi == sent).all()), "%s: write outside the buffer" % nm
