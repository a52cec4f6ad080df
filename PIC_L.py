"""Drop-in replacement for the reference's PIC_L.py (periodic explicit leapfrog PIC with a
Poisson solve every step, plus the function forms of the implicit push): same module-level
names, signatures and return values, backed by the sm_100a CUDA kernels of libpic_b200.so
through pypic_b200.  No CPU fallback: every numerical function runs on the GPU; only
configuration, the legacy-RNG initialiser and I/O stay on the host.

The reference file is Python 2 (``N/2`` used as an index); the integer divisions are
written ``//`` here.  Functions the reference's drivers never call are served by device kernels too; only main_i
(a duplicate of pypic.implicit_pic with a stale-variable bug) is not rebuilt.
"""
from __future__ import print_function

import ctypes as C

import numpy as np

from pypic_b200 import ops
from pypic_b200.periodic import ExplicitSim
from pypic_b200.plotting import get_plt

np.random.seed(1)          # PIC_L.py:20 seeds the global legacy stream at import

lw = 3.0

# physical constants (PIC_L.py:24-30)
epsilon0 = 8.854E-12
e = 1.602E-19
mp = 1.67E-27
me = 9.11E-31
kb = 1.38E-23


def _scalar_or_array(out, x):
    return float(out[0]) if np.ndim(x) == 0 else out


def interpolateField(F, x, dx):
    """PIC_L.py:32-37 (bounded gather; same arithmetic as PIC_L_DD.interpolateField)."""
    F = np.asarray(F, dtype=np.float64)
    return _scalar_or_array(ops.dd_interpolate(F, x, len(F), dx), x)


def interpolateFieldPeriodic(F, x, Ng, dx):
    """PIC_L.py:39-46.  Accepts a scalar (like the reference) or an array of positions."""
    return _scalar_or_array(ops.l_interpolate(F, x, Ng, dx), x)


def weightCurrents(x, q, v, p2c, Ng, N, dx):
    """PIC_L.py:48-60: bounded CIC current on Ng nodes, without wall terms or edge fold."""
    return ops.l_weight_bounded(np.asarray(x)[:N], np.asarray(q)[:N], np.asarray(v)[:N], p2c, Ng, dx)


def weightCurrentsPeriodic(x, q, v, p2c, Ng, N, dx):
    """PIC_L.py:62-80."""
    return ops.l_weight(np.asarray(x)[:N], np.asarray(q)[:N], np.asarray(v)[:N], p2c, Ng, dx)


def weightDensities(x, q, p2c, Ng, N, dx):
    """PIC_L.py:83-98: bounded CIC density on Ng nodes."""
    return ops.l_weight_bounded(np.asarray(x)[:N], np.asarray(q)[:N], None, p2c, Ng, dx)


def weightDensitiesPeriodic(x, q, p2c, Ng, N, dx):
    """PIC_L.py:100-118."""
    return ops.l_weight(np.asarray(x)[:N], np.asarray(q)[:N], None, p2c, Ng, dx)


def laplacian1DPeriodic(Ng):
    """PIC_L.py:120-132 (matrix constructor over the Ng+1 nodes; host-side helper)."""
    A = np.diag(np.ones(Ng), -1) + np.diag(-2. * np.ones(Ng + 1), 0) + np.diag(np.ones(Ng), 1)
    A[0, -1] = 1.
    A[-1, 0] = 1.
    return A


def laplacian1D(Ng):
    """PIC_L.py:134-144."""
    A = np.diag(np.ones(Ng - 1), -1) + np.diag(-2. * np.ones(Ng), 0) + np.diag(np.ones(Ng - 1), 1)
    A[0, 0] = 1.
    A[0, 1] = 0.
    A[0, 2] = 0.
    A[-1, -1] = -2.
    A[-1, -2] = 1.
    A[-1, -3] = 1.
    return A


def solvePoisson(dx, Ng, rho, kBT, tol, maxiter, phi0):
    """PIC_L.py:146-177: Boltzmann-Newton solve on Ng nodes as written (reference node Ng/2, the
    three-entry last row of laplacian1D, `while resid > tol and k <= maxiter`), whole loop in one
    kernel launch with a PCR solve per iteration in place of scipy.sparse.linalg.inv."""
    return ops.newton_boltzmann_l(np.asarray(rho)[:Ng], phi0, dx, kBT, tol, maxiter, periodic=False)


def solvePoissonPeriodic(dx, Ng, rho, kBT, tol, maxiter, phi0):
    """PIC_L.py:179-206: the periodic Boltzmann-Newton solve on the Ng+1 nodes (cyclic tridiagonal
    system by Sherman-Morrison over PCR)."""
    return ops.newton_boltzmann_l(np.asarray(rho)[:Ng + 1], phi0, dx, kBT, tol, maxiter, periodic=True)


def solvePoissonPeriodicElectronsNeutralized(dx, Ng, rho, kBT, tol, maxiter, phi0):
    """PIC_L.py:208-220 over the Ng+1 nodes.  The periodic matrix is singular; the gauge is
    phi[-1]=0 (callers subtract max(phi), PIC_L.py:690,765)."""
    return ops.poisson_periodic(rho, dx, subtract_max=False)


def differentiateField(F, dx, Ng):
    """PIC_L.py:222-233 (same arithmetic as PIC_L_DD.differentiateField)."""
    return ops.differentiate(F, dx, 1)


def differentiateFieldPeriodic(F, dx, Ng):
    """PIC_L.py:235-246."""
    return ops.differentiate(F, dx, 2)


def _uniform(a, name):
    a = np.asarray(a, dtype=np.float64)
    if a.size and np.any(a != a.flat[0]):
        return None
    return float(a.flat[0]) if a.size else 0.0


def pushParticlesExplicit(x, v, q, m, N, Ng, dt, dx, E):
    """PIC_L.py:248-259: periodic gather + kick-drift-kick; returns the UNWRAPPED xout, vout."""
    import torch
    from pypic_b200 import _lib, device as D
    x = np.asarray(x, dtype=np.float64)[:N]; v = np.asarray(v, dtype=np.float64)[:N]
    q = np.asarray(q, dtype=np.float64)[:N]; m = np.asarray(m, dtype=np.float64)[:N]
    # the device store keeps one (q, m) pair per species block: split at the first change
    change = np.nonzero((q[1:] != q[:-1]) | (m[1:] != m[:-1]))[0]
    if len(change) > 1:
        raise NotImplementedError("more than two contiguous species blocks in pushParticlesExplicit")
    ns = int(change[0]) + 1 if len(change) else N
    q2 = (float(q[0]), float(q[ns] if ns < N else q[0])) if N else (0., 0.)
    m2 = (float(m[0]), float(m[ns] if ns < N else m[0])) if N else (1., 1.)
    dev = D.require_cuda()
    L = dx * (Ng - 1)
    P = _lib.LParams(N, ns, Ng, 2, dx, dt, L, 1.0, (C.c_double * 2)(*q2), (C.c_double * 2)(*m2))
    tx, tv, tE = D.to_dev(x, dev), D.to_dev(v, dev), D.to_dev(np.asarray(E, dtype=np.float64), dev)
    acc = D.f64(Ng + 1, dev, True)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("pic_dev_l_push_deposit", C.byref(P), D.ptr(tx), D.ptr(tv), D.ptr(tE), D.ptr(acc), D.ptr(err), D.stream())
    D.check_range(err, "PIC_L.pushParticlesExplicit")
    return tx.cpu().numpy(), tv.cpu().numpy()


def pushParticlesImplicit(x0, xh, v, q, m, N, Ng, dt, dx, Eh):
    """PIC_L.py:261-270: function form of the Crank-Nicolson push (periodic gather of Eh at xh)."""
    f = lambda a: np.asarray(a, dtype=np.float64)[:N]
    return ops.l_push_implicit(f(x0), f(xh), f(v), f(q), f(m), Ng, dt, dx, Eh)


def applyBoundaryConditions(x, v, m, N, L, dx, kBT):
    """PIC_L.py:272-282: particles with x > L or x <= 0 are redrawn -- x = uniform(dx, L-dx),
    v = normal(0, sqrt(kBT/mp)) -- in index order from the global legacy stream.  The device finds
    them (flag + stable compaction); the draws are the reference's; x and v are modified in place
    and returned, like there."""
    idx = ops.l_outside(np.asarray(x, dtype=np.float64)[:N], L)
    s = np.sqrt(kBT / mp)
    for i in idx:
        x[i] = np.random.uniform(1. * dx, L - 1. * dx)
        v[i] = np.random.normal(0.0, s, 1)[0]
    return x, v


def applyBoundaryConditionsPeriodic(x, v, m, N, L, dx, kBT):
    """PIC_L.py:284-288: x % (L+dx)."""
    import torch
    from pypic_b200 import _lib, device as D
    dev = D.require_cuda()
    tx = D.to_dev(np.asarray(x, dtype=np.float64)[:N], dev)
    _lib.call("pic_dev_wrap_periodic", D.ptr(tx), tx.numel(), float(L + dx), D.stream())
    return tx.cpu().numpy(), v


def initialize(system, N, density, Kp, perturbation, dx, Ng, Te, L, X):
    """PIC_L.py:290-366.  Host-side initialiser drawing from the global legacy np.random
    stream in the reference's order."""
    wp = np.sqrt(e**2 * density / epsilon0 / me)
    K = Kp * np.pi / (L + dx)
    kBTe = kb * Te
    vthermal = np.sqrt(2.0 * kBTe / me)
    LD = 7430.0 * np.sqrt(kBTe / e / density)
    print('debye length * kp: ', LD * K)
    m = np.ones(N) * me
    q = -np.ones(N) * e
    if system == 'bump-on-tail':
        beam, plasma = N * 2 // 6, N * 4 // 6
        growth_rate = np.sqrt(3.) / 2. * wp * (float(beam) / float(plasma) / 2.)**(1. / 3.)
        v0 = np.zeros(N)
        v0[0:plasma] = np.random.normal(0.0, np.sqrt(kBTe / me), plasma)
        v0[plasma:] = np.random.normal(5.0 * np.sqrt(kBTe / me), (1. / 20.) * np.sqrt(kBTe / me), beam + 1)
    elif system == 'landau damping':
        growth_rate = -np.sqrt(np.pi) * wp * (wp / K / vthermal)**3 * np.exp(-wp**2 / K**2 / vthermal**2) * np.exp(-3. / 2.)
        v0 = np.random.normal(0.0, np.sqrt(kBTe / me), N)
    elif system == 'two-stream':
        b1 = N // 2
        b2 = N - b1
        v0 = np.zeros(N)
        v0[0:b1] = np.random.normal(1.0 * np.sqrt(kBTe / me), (1. / 40.) * np.sqrt(kBTe / me), b1)
        v0[b1:] = np.random.normal(0.0, (1. / 40.) * np.sqrt(kBTe / mp), b2)
        m[b1:] = mp
        growth_rate = wp * (me / mp)**(1. / 3.)
    else:
        raise ValueError("unknown system %r" % (system,))
    x0 = np.random.uniform(0., L + dx, N)
    F = -np.cos(Kp * np.pi * X / (L + dx)) + 1.0
    F = (N * perturbation) * F / np.sum(F)
    j = N // 2 - int(N * perturbation / 2)
    for i in range(Ng):
        for k in range(int(F[i])):
            x0[j] = np.random.uniform(X[i], X[i + 1])
            j += 1
    x0 = x0 % (L + dx)
    return m, q, x0, v0, kBTe, growth_rate


def main_i(T, nplot):
    raise NotImplementedError("PIC_L.main_i (PIC_L.py:368-602) duplicates pypic.implicit_pic and gathers with stale "
                              "variables (Eh,xh instead of Es,xs, PIC_L.py:486-488); the implicit periodic loop on the "
                              "GPU is pypic.main / pypic.implicit_pic")


def main(T, nplot, system='landau damping', density=1e10, perturbation=0.05, Kp=2, N=100000, Ng=200, dt=1E-9, dx=0.02,
         Te=10.0 * 11600., outdir='plots', result=None, sort_every=None, deposit='warp'):
    """PIC_L.main (PIC_L.py:604-786): explicit leapfrog loop, Poisson solve every step.  The
    positional signature is the reference's; the keyword arguments default to its hard-coded
    literals.  Particles stay resident on the GPU; `result` (a dict) receives the series.
    The store is re-sorted by (species, cell) every `sort_every` steps (default: 16 from 2^17 particles
    on, never below) for the window kernel; downloads return the reference's particle order.
    deposit='window-det': the reproducible build (ExplicitSim) -- two runs give bit-identical output."""
    L = dx * (Ng - 1)
    X = np.linspace(0.0, L + dx, Ng + 1)
    wp = np.sqrt(e**2 * density / epsilon0 / me)
    invwp = 1. / wp
    K = Kp * np.pi / (L + dx)
    p2c = (L + dx) * density / N
    m, q, x0, v0, kBTe, growth_rate = initialize(system, N, density, Kp, perturbation, dx, Ng, Te, L, X)
    print("wp : ", wp, "[1/s]")
    print("dt : ", dt / invwp, " [w * tau]")
    print("tau: ", invwp, "[s]")
    print("k  : ", K, "[1/m]")
    print("p2c :", p2c)
    change = np.nonzero(m[1:] != m[:-1])[0]
    ns = int(change[0]) + 1 if len(change) else N
    if sort_every is None:
        sort_every = 16 if N >= (1 << 17) else 0
    sim = ExplicitSim(N, Ng, dx, dt, p2c, q=(float(q[0]), float(q[-1])), m=(float(m[0]), float(m[-1])), n_split=ns,
                      sort_every=sort_every, deposit=deposit)
    sim.upload(x0, v0)
    mpl, plt = get_plt()
    EE, KE, TT, E_series = [], [], [], []
    sim.field_solve()                               # PIC_L.py:688-691
    for t in range(T):
        print('t: ', t)
        EE.append(sim.field_energy())               # :697  sum(eps0 E^2/2) of the field in use
        KE.append(sim.kinetic_energy(me))           # :698  sum(me v^2/2)
        TT.append(dt * t)
        if result is not None:
            E_series.append(sim.E.cpu().numpy())
        if plt is not None and (t % nplot == 0):
            st = sim.download()
            plt.figure(1); plt.clf()
            plt.scatter(st["x"], st["v"] / np.sqrt(kBTe / me), s=0.5)
            plt.savefig(outdir + '/ps_' + str(t))
            plt.figure(3); plt.clf()
            plt.plot(X, st["E"], linewidth=lw)
            plt.savefig(outdir + '/e_' + str(t))
        sim.step()                                  # :763-768
    sim.check()
    np.savetxt(outdir + '/E2.txt', EE)
    np.savetxt(outdir + '/J.txt', np.zeros(Ng + 1))         # the reference saves its never-updated j (PIC_L.py:668,771)
    with open(outdir + '/parameters.out', 'w+') as output_file:
        for name, val in (('wp', wp), ('Te', Te), ('G', growth_rate), ('tau', 1.0 / wp), ('p2c', p2c), ('dt', dt),
                          ('dx', dx), ('Ng', Ng), ('L', L + dx)):
            print(name, val, file=output_file)
    if result is not None:
        result.update(sim.download(), EE=np.array(EE), KE=np.array(KE), TT=np.array(TT), E_series=np.array(E_series))
    return EE


if __name__ == '__main__':
    main(200, 10)
