"""Reproducible build of the sheath path (`SheathSim(deposit="window-det")`): order-independent
fixed-point accumulation of the deposits + stable radix sort by cell.  North star (2): deposition
with a deterministic ordering next to the fast atomic variant."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import np_oracle as O


def _sim(N, Ng, deposit, sort_every, seed=99, **kw):
    import torch
    from pypic_b200 import device as D
    from pypic_b200.sheath import SheathSim
    dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1)
    kT = O.kb * 116000.
    dev = D.require_cuda()
    s = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=False, deposit=deposit, rng="philox", seed=1,
                  device=dev, sort_every=sort_every, **kw)
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    s.x0.uniform_(0.0, 1.0, generator=gen).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
    s.u0.normal_(0.0, 1.0, generator=gen)
    s.u0[:s.n_split].mul_(float(np.sqrt(kT / O.me))); s.u0[s.n_split:].mul_(float(np.sqrt(kT / O.mp)))
    s.E0.normal_(0.0, 1e4, generator=gen)
    return s


@pytest.mark.parametrize("N,Ng,n_split", [(300001, 4097, None), (70000, 51, 1234), (8192, 257, 0), (123457, 70001, 60000),
                                          (4096, 300, 4096), (1, 51, 1)])
def test_stable_sort_is_numpy_stable_argsort(N, Ng, n_split):
    """pic_dev_dd_sort_by_cell_stable == np.argsort(cell, kind='stable') inside each species block
    (1, 2 and 3 radix passes; ragged tiles; empty and one-particle blocks)."""
    import torch
    from pypic_b200 import _lib, device as D
    dev = D.require_cuda()
    rs = np.random.RandomState(N % 1000)
    dx = 1e-5; L = dx * (Ng - 1)
    x = rs.uniform(0, L, N)
    x[: min(N, Ng - 1)] = np.arange(min(N, Ng - 1)) * dx          # node-aligned positions
    rs.shuffle(x)
    u = rs.normal(0, 1, N)
    ns = N // 2 if n_split is None else n_split
    P = _lib.DDParams(N, ns, Ng, 128, dx, 1e-12, L, 1.0, (C.c_double * 2)(-O.e, O.e), (C.c_double * 2)(O.me, O.mp))
    dx0, du0 = D.to_dev(x, dev), D.to_dev(u, dev)
    xs, us = torch.full_like(dx0, -1.0), torch.full_like(du0, -1.0)
    scratch = torch.zeros(D.sort_stable_scratch_size(N), dtype=torch.int32, device=dev)
    where = C.c_int(-1)
    _lib.call("pic_dev_dd_sort_by_cell_stable", C.byref(P), D.ptr(dx0), D.ptr(du0), D.ptr(xs), D.ptr(us),
              D.ptr(scratch), scratch.numel(), C.byref(where), D.stream())
    gx, gu = (xs, us) if where.value else (dx0, du0)
    gx, gu = gx.cpu().numpy(), gu.cpu().numpy()
    cell = np.clip(np.floor(x / dx).astype(np.int64), 0, Ng - 1)
    for lo, hi in ((0, ns), (ns, N)):
        o = lo + np.argsort(cell[lo:hi], kind="stable")
        assert np.array_equal(gx[lo:hi], x[o]) and np.array_equal(gu[lo:hi], u[o])


@pytest.mark.parametrize("sort", [False, True])
def test_fixed_point_deposit_is_schedule_independent_and_matches_fp64_atomics(sort):
    """One Picard iteration (first and later) of the reproducible build: particles bit-identical to
    the default build; the accumulated fixed-point WORDS bit-identical between the dynamically
    scheduled and the static (flags bit6) launch and between repeated launches; currents equal to
    the fp64-atomic ones to round-off.  Unsorted: every deposit takes the per-particle global path;
    sorted: the private windows."""
    import torch
    from pypic_b200 import _lib, device as D
    N, Ng = 148 * 16384 * 2 + 12345, 1025
    ref = _sim(N, Ng, "window", 8)
    det = _sim(N, Ng, "window-det", 8)
    if sort:
        det.sort_by_cell()
        ref.x0.copy_(det.x0); ref.u0.copy_(det.u0)
    assert torch.equal(ref.x0, det.x0)
    for s in (ref, det):
        s.Es.copy_(s.E0)
    g = Ng
    for it in range(3):
        words = []
        x1_in, act_in = det.x1.clone(), det.active.clone()      # what a non-first iteration reads
        for flags in (128, 128 | 64, 128):
            det.params.flags = flags
            det.acc.zero_(); det.x1.copy_(x1_in); det.active.copy_(act_in)
            _lib.call("pic_dev_dd_picard_iter", C.byref(det.params), D.ptr(det.x0), D.ptr(det.u0), D.ptr(det.x1),
                      D.ptr(det.u1), D.ptr(det.active), D.ptr(det.Es), D.ptr(det.acc), 1 if it == 0 else 0,
                      D.ptr(det.range_err), D.stream())
            words.append(det.acc.clone())
        assert torch.equal(words[0].view(torch.int64), words[1].view(torch.int64)), it
        assert torch.equal(words[0].view(torch.int64), words[2].view(torch.int64)), it
        assert float(words[0][:2 * g].abs().max()) == 0.0          # nothing went to the fp64 slots
        ref.acc.zero_()
        _lib.call("pic_dev_dd_picard_iter", C.byref(ref.params), D.ptr(ref.x0), D.ptr(ref.u0), D.ptr(ref.x1),
                  D.ptr(ref.u1), D.ptr(ref.active), D.ptr(ref.Es), D.ptr(ref.acc), 1 if it == 0 else 0,
                  D.ptr(ref.range_err), D.stream())
        assert torch.equal(ref.x1, det.x1) and torch.equal(ref.u1, det.u1) and torch.equal(ref.active, det.active)
        assert torch.equal(ref.acc[2 * g:2 * g + 4], det.acc[2 * g:2 * g + 4])          # absorbed counts
        w = det.acc[2 * g + 4:].view(torch.int64).cpu().numpy()
        from pypic_b200 import fixedpoint as FP
        sc = FP.scale_exponent(det.q, det.p2c, det.dx)
        cur = np.array([FP.merge(int(w[n]), int(w[2 * g + n]), sc) for n in range(2 * g)])
        a = ref.acc[:2 * g].cpu().numpy()
        assert np.abs(cur - a).max() < 1e-12 * np.abs(a).max(), it
        ref.Es.mul_(0.97); det.Es.copy_(ref.Es)
    assert int((det.active != 1).sum()) > 0
    ref.check(); det.check()


def test_reproducible_run_is_bit_identical_twice_and_close_to_default():
    """Whole steps (re-injection, stable sort every 2 steps, light Picard iterations, repair
    passes): two runs of the reproducible build end in bit-identical particles, flags and fields;
    a third run on the static schedule too.  Against the default build (no sort, so that the
    Philox draws, keyed by slot, stay comparable) the fields agree to round-off."""
    import torch
    N, Ng = 148 * 16384 + 5000, 513
    outs = []
    for blocked in (False, False, True):
        s = _sim(N, Ng, "window-det", 2)
        if blocked:
            s.params.flags |= 64
        its = [s.step()[0] for _ in range(6)]
        s.check()
        outs.append((its, s.x0.clone(), s.u0.clone(), s.active.clone(), s.E0.clone(), s.j0.clone()))
    for o in outs[1:]:
        assert o[0] == outs[0][0]
        for a, b in zip(o[1:], outs[0][1:]):
            assert torch.equal(a, b)
    assert int((outs[0][3] != 1).sum()) > 0          # walls absorbed something in the last step
    a = _sim(N, Ng, "window-det", 0)
    b = _sim(N, Ng, "window", 0)
    for _ in range(3):
        ka = a.step()[0]; kb = b.step()[0]
        assert ka == kb
    Ea, Eb = a.E0.cpu().numpy(), b.E0.cpu().numpy()
    assert np.abs(Ea - Eb).max() < 1e-10 * np.abs(Eb).max()
    assert torch.equal(a.active, b.active)


def test_reproducible_run_on_a_grid_beyond_one_cta():
    """Ng > 32768: the field phase is the cooperative multi-CTA kernel, whose four global sums are
    formed from per-CTA partials added in CTA order -- two runs of the reproducible build (large-grid
    particle kernel, stable sort, fixed-point deposits) are bit-identical there too."""
    import torch
    N, Ng = 40 * 16384 + 321, 70001
    outs = []
    for _ in range(2):
        s = _sim(N, Ng, "window-det", 2)
        its = [s.step()[0] for _ in range(4)]
        s.check()
        outs.append((its, s.x0.clone(), s.u0.clone(), s.active.clone(), s.E0.clone(), s.j0.clone()))
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert torch.equal(a, b)
    d = _sim(N, Ng, "window", 0); r = _sim(N, Ng, "window-det", 0)
    for _ in range(2):
        assert d.step()[0] == r.step()[0]
    Ed, Er = d.E0.cpu().numpy(), r.E0.cpu().numpy()
    assert np.abs(Ed - Er).max() < 1e-10 * np.abs(Ed).max()


@pytest.mark.parametrize("N,Ng,n_split", [(300001, 4097, None), (70000, 51, 1234), (123457, 70001, 60000), (4096, 300, 4096)])
def test_stable_sort_carries_the_original_index(N, Ng, n_split):
    """pic_dev_dd_sort_by_cell_stable2: the int32 payload travels with (x, u) through every radix pass -- first sort
    from the identity numbering, second sort (after the particles moved) from the carried one: the payload is
    np.argsort(cell, kind='stable') composed with the previous permutation."""
    import torch
    from pypic_b200 import _lib, device as D
    dev = D.require_cuda()
    rs = np.random.RandomState(N % 997)
    dx = 1e-5; L = dx * (Ng - 1)
    x = rs.uniform(0, L, N); u = rs.normal(0, 1, N)
    ns = N // 2 if n_split is None else n_split
    P = _lib.DDParams(N, ns, Ng, 128, dx, 1e-12, L, 1.0, (C.c_double * 2)(-O.e, O.e), (C.c_double * 2)(O.me, O.mp))
    bufs = [D.to_dev(x, dev), D.to_dev(u, dev), torch.empty(N, dtype=torch.float64, device=dev),
            torch.empty(N, dtype=torch.float64, device=dev)]
    oid = [torch.full((N,), -7, dtype=torch.int32, device=dev), torch.full((N,), -7, dtype=torch.int32, device=dev)]
    scratch = torch.zeros(D.sort_stable_scratch_size(N), dtype=torch.int32, device=dev)
    cur_x, cur_u, cur_o = x.copy(), u.copy(), np.arange(N)
    for rnd in range(2):
        where = C.c_int(-1)
        _lib.call("pic_dev_dd_sort_by_cell_stable2", C.byref(P), D.ptr(bufs[0]), D.ptr(bufs[1]), D.ptr(bufs[2]), D.ptr(bufs[3]),
                  D.ptr(oid[0]), D.ptr(oid[1]), 1 if rnd == 0 else 0, D.ptr(scratch), scratch.numel(), C.byref(where), D.stream())
        if where.value:
            bufs = [bufs[2], bufs[3], bufs[0], bufs[1]]; oid = [oid[1], oid[0]]
        cell = np.clip(np.floor(cur_x / dx).astype(np.int64), 0, Ng - 1)
        o = np.concatenate([lo + np.argsort(cell[lo:hi], kind="stable") for lo, hi in ((0, ns), (ns, N))]).astype(np.int64)
        cur_x, cur_u, cur_o = cur_x[o], cur_u[o], cur_o[o]
        assert np.array_equal(bufs[0].cpu().numpy(), cur_x) and np.array_equal(bufs[1].cpu().numpy(), cur_u)
        assert np.array_equal(oid[0].cpu().numpy().astype(np.int64), cur_o)
        # the particles move before the next sort
        cur_x = np.clip(cur_x + rs.normal(0, 3 * dx, N), 0, L * (1 - 1e-12))
        bufs[0].copy_(torch.as_tensor(cur_x))


def test_reproducible_run_through_the_reference_api_keeps_the_particle_numbering(tmp_path):
    """PIC_L_DD.main_i(deposit='window-det'): the reference's own loop (legacy MT19937 re-injection in index order,
    carried v,w) on the reproducible build WITH the cell sort -- the stable radix sort carries the original index.
    Two runs give bit-identical series and particles (in the reference's numbering); against the default build
    the run agrees to round-off."""
    import contextlib, io, os
    import PIC_L_DD
    outs = []
    for dep in ("window-det", "window-det", "window"):
        res = {}
        np.random.seed(3)
        cwd = os.getcwd(); os.chdir(str(tmp_path))
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                PIC_L_DD.main_i(9, 10 ** 9, N=120000, Ng=257, result=res, sort_every=2, deposit=dep)
        finally:
            os.chdir(cwd)
        outs.append(res)
    a, b, d = outs
    keys = [k for k in a if isinstance(a[k], np.ndarray)]
    assert {"x0", "u0", "E0"} <= set(keys)
    for k in keys:
        assert np.array_equal(a[k], b[k]), k                      # bit for bit, diagnostics included
    for k in ("x0", "u0", "E0"):
        assert np.max(np.abs(a[k] - d[k])) <= 1e-9 * np.max(np.abs(d[k])), k


@pytest.mark.parametrize("N", [200000, 6000])       # 6000: shorter than one chunk, every particle through the exact routine
def test_explicit_loop_reproducible_build_through_the_reference_api(tmp_path, N):
    """PIC_L.main(deposit='window-det'): the explicit leapfrog loop on the reproducible build of its window kernel
    (fixed-point merges into rho, stable radix sort with the original-index payload, fixed-order kinetic-energy sum).
    Two runs give bit-identical series, fields and particles (in the reference's numbering); against the default
    build the run agrees to round-off."""
    import contextlib, io, os
    import PIC_L
    outs = []
    os.makedirs(os.path.join(str(tmp_path), "plots"), exist_ok=True)
    for dep, se in (("window-det", 3), ("window-det", 3), ("warp", 0)):
        res = {}
        np.random.seed(5)
        cwd = os.getcwd(); os.chdir(str(tmp_path))
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                PIC_L.main(8, 10 ** 9, N=N, result=res, sort_every=se, deposit=dep)
        finally:
            os.chdir(cwd)
        outs.append(res)
    a, b, d = outs
    keys = [k for k in a if isinstance(a[k], np.ndarray)]
    assert {"x", "v", "E", "EE", "KE", "E_series"} <= set(keys)
    for k in keys:
        assert np.array_equal(a[k], b[k]), k
    for k in ("x", "v", "E", "EE", "E_series"):
        assert np.max(np.abs(a[k] - d[k])) <= 1e-9 * np.max(np.abs(d[k])), k
    assert np.max(np.abs(a["KE"] - d["KE"])) <= 1e-12 * np.max(np.abs(d["KE"]))


@pytest.mark.parametrize("N", [200000, 9000])       # 9000: shorter than one chunk, every particle through the exact routine
def test_periodic_implicit_loop_reproducible_build_through_the_reference_api(tmp_path, N):
    """pypic.main(deposit='window-det'): the periodic Crank-Nicolson / Picard loop on the reproducible build of its
    window kernel (fixed-point merges of jh and j1, also in the repair pass of a loop that ended on a light iteration;
    order-independent initial density; stable radix sort with the original-index payload; fixed-order kinetic
    energy).  Two runs give bit-identical series, fields and particles (in the reference's numbering); against
    the default build the run agrees to round-off."""
    import contextlib, io, os
    import pypic
    outs = []
    os.makedirs(os.path.join(str(tmp_path), "plots"), exist_ok=True)
    for dep, se in (("window-det", 2), ("window-det", 2), ("warp", 0)):
        res = {}
        np.random.seed(6)
        cwd = os.getcwd(); os.chdir(str(tmp_path))
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                pypic.main(7, 10 ** 9, N=N, Ng=64, result=res, sort_every=se, deposit=dep)
        finally:
            os.chdir(cwd)
        outs.append(res)
    a, b, d = outs
    keys = [k for k in a if isinstance(a[k], np.ndarray)]
    assert {"x0", "v0", "E0", "j0", "EE", "KE", "j_bias"} <= set(keys)
    for k in keys:
        assert np.array_equal(a[k], b[k]), k
    for k in ("x0", "v0", "E0", "EE"):
        assert np.max(np.abs(a[k] - d[k])) <= 1e-9 * np.max(np.abs(d[k])), k
    assert np.max(np.abs(a["KE"] - d["KE"])) <= 1e-12 * np.max(np.abs(d["KE"]))
