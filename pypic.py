"""Drop-in replacement for the reference's pypic.py (periodic implicit Crank-Nicolson /
Picard PIC): same module-level names, signatures and outputs, backed by the sm_100a CUDA
kernels of libpic_b200.so through pypic_b200.  `run_pypic.py` of the reference drives this
module unchanged (tools/drive.py).  No CPU fallback.

Reference quirks kept on purpose: `p2c` is truncated to an integer in
the kernels that numba types as int32; `main` passes `Nv = Ng/2` (a float) which makes
NumPy 2 raise inside the reference -- here it is made an int.
"""
from __future__ import print_function

import numpy as np

from pypic_b200 import ops
from pypic_b200.periodic import PeriodicImplicitSim
from pypic_b200.plotting import get_plt

lw = 3.0

# physical constants (pypic.py:22-26)
epsilon0 = 8.854E-12
e = 1.602E-19
mp = 1.67E-27
me = 9.11E-31
kb = 1.38E-23


def interpolate_p(F, x, Ng, N, dx):
    """pypic.py:28-61."""
    return ops.pypic_interpolate(F, np.asarray(x)[:N], Ng, dx)


def smooth_field_p(F):
    """pypic.py:63-76."""
    return ops.smooth(F, 0)


def find_cell_indices_and_weights_p(x, Ng, N, dx):
    """pypic.py:78-89 (unused by the reference; its decorator types dx as an array)."""
    raise NotImplementedError("find_cell_indices_and_weights_p is dead code in the reference (never called)")


def weight_current_p(x, q, v, p2c, Ng, N, dx):
    """pypic.py:91-136 (p2c truncated like numba's int32 argument)."""
    return ops.pypic_weight(np.asarray(x)[:N], np.asarray(q)[:N], np.asarray(v)[:N], p2c, Ng, dx)


def weight_density_p(x, q, p2c, Ng, N, dx):
    """pypic.py:138-183."""
    return ops.pypic_weight(np.asarray(x)[:N], np.asarray(q)[:N], None, p2c, Ng, dx)


def differentiate_p(F, dx, Ng):
    """pypic.py:185-214."""
    return ops.differentiate(F, dx, 0)


def _scalar_or_array(a):
    """One species (every entry equal: the reference's only use) -> scalar, which takes the fused TMA
    kernel; genuinely per-particle values -> the array, served by the grid-stride kernel."""
    a = np.asarray(a, dtype=np.float64)
    if a.size == 0 or not np.any(a != a.flat[0]):
        return float(a.flat[0]) if a.size else 0.0
    return a


def particle_push_p(x0, v0, q, m, E0, j0, N, Ng, p2c, dx, dt, L, tol, maxiter):
    """pypic.py:216-300: implicit particle push + field advance.  Returns x1, v1, E1, j1."""
    sim = PeriodicImplicitSim(N, Ng, dx, dt, L, p2c, q=_scalar_or_array(q), m=_scalar_or_array(m), tol=tol, maxiter=maxiter)
    sim.upload(np.asarray(x0, dtype=np.float64), np.asarray(v0, dtype=np.float64), np.asarray(E0, dtype=np.float64))
    k, r = sim.push()
    sim.check()
    print("Iterations: ", k)
    print("Residual  : ", r)
    o = sim.download()
    if k == 0:
        return np.array(x0), np.array(v0), np.array(E0), np.array(j0)
    return o["x0"], o["v0"], o["E0"], o["j0"]


def differentiate_t(F, dt):
    """pypic.py:302-335 (host-side post-processing of the energy time series)."""
    F = np.asarray(F, dtype=np.float64)
    T = len(F)
    dF = np.zeros(T)
    dF[0] = (F[1] - F[0]) / dt
    dF[1:T - 1] = (F[2:] - F[:T - 2]) / dt * 0.5
    dF[T - 1] = (F[T - 1] - F[T - 2]) / dt
    return dF


def laplacian_1D_p(Ng):
    """pypic.py:337-357 (matrix constructor; host-side helper)."""
    A = np.diag(np.ones(Ng - 1), -1) + np.diag(-2. * np.ones(Ng), 0) + np.diag(np.ones(Ng - 1), 1)
    A[0, -1] = 1.
    A[-1, 0] = 1.
    return A


def solve_poisson_p(dx, Ng, rho, phi0):
    """pypic.py:359-382: periodic Poisson solve (PCR tridiagonal on the device; the singular
    system is gauge-fixed with phi[-1] = 0, callers subtract max(phi))."""
    return ops.poisson_periodic(rho, dx, subtract_max=False)


def initialize_p(system, N, density, Kp, perturbation, dx, Ng, Te, Ti, L, X):
    """pypic.py:384-470.  Host-side initialiser drawing from the global legacy np.random
    stream in the reference's order (velocities, positions, then the per-cell resampling)."""
    wp = np.sqrt(e**2 * density / epsilon0 / me)
    invwp = 1. / wp
    K = Kp * 2.0 * np.pi / L
    p2c = L * density / N
    kBTe = kb * Te
    kBTi = kb * Ti
    v_thermal = np.sqrt(2.0 * kBTe / me)
    LD = np.sqrt(kBTe * epsilon0 / e / e / density)
    m = np.ones(N) * me
    q = -np.ones(N) * e
    vt = np.sqrt(kBTe / me)
    if system == 'bump-on-tail':
        nb, npl = N * 1 // 6, N * 5 // 6
        growth_rate = np.sqrt(3.) / 2. * wp * (float(nb) / float(npl) / 2.)**(1. / 3.)
        v0 = np.zeros(N)
        v0[0:npl] = np.random.normal(0.0, vt, npl)
        v0[npl:] = np.random.normal(4.0 * vt, (1. / 20.) * vt, nb + 1)
    elif system == 'two-stream':
        n1 = n2 = N * 1 // 2
        growth_rate = np.sqrt(3.) / 2. * wp * (float(n1) / float(n2) / 2.)**(1. / 3.)
        v0 = np.zeros(N)
        v0[0:n1] = np.random.normal(-2.0 * vt, 0.5 * vt, n1)
        v0[n1:] = np.random.normal(2.0 * vt, 0.5 * vt, n2)
    elif system == 'landau-damping':
        v0 = np.random.normal(0.0, v_thermal / np.sqrt(2), N)
        growth_rate = -np.sqrt(np.pi) * wp * (wp / K / v_thermal)**3 * np.exp(-1. / (2.0 * K**2 * LD**2) - 3. / 2.)
    else:
        raise ValueError("unknown system %r" % (system,))
    x0 = np.random.uniform(0.0, L, N)
    F = 1.0 + np.cos(K * X)
    F = (N * perturbation) * F / np.sum(F)
    j = 0
    for i in range(Ng):
        for k in range(int(F[i])):
            x0[j] = np.random.uniform(X[i], X[i + 1])
            j += 1
    return m, q, x0, v0, kBTe, kBTi, growth_rate, K, p2c, wp, invwp, LD


def implicit_pic(T, nplot, system, density, perturbation, Kp, N, Ng, Nv, Vmax, dt, Ti, Te, L, tol, maxiter,
                 outdir='plots', result=None, sort_every=None, deposit='warp'):
    """pypic.py:472-651: main implicit PIC routine (particles stay resident on the GPU).
    deposit='window-det': the reproducible build (PeriodicImplicitSim) -- two runs give bit-identical output.

    The store is re-sorted by cell every `sort_every` steps (default: 8 from 2^17 particles on, never below) so that
    the loop runs on the TMA-staged window kernel; the original index of every particle rides along, so the
    tracer, the plots and `result` see the reference's particle order."""
    tracer = 9999
    Nv = int(Nv)
    X = np.linspace(0.0, L, Ng + 1)
    dx = L / float(Ng)
    m, q, x0, v0, kBTe, kBTi, growth_rate, K, p2c, wp, invwp, LD = initialize_p(system, N, density, Kp, perturbation,
                                                                                dx, Ng, Te, Ti, L, X)
    print("wp : ", wp, "[1/s]")
    print("dt : ", dt / invwp, " [dt * wp]")
    print("tau: ", invwp, "[s]")
    print("k*LD: ", K * LD)
    print("p2c :", p2c)
    print("gamma: ", growth_rate)
    KE, EE, TT, j_bias, trajectory_x, trajectory_v = [], [], [], [], [], []
    # initial field from one Poisson solve (pypic.py:550-554)
    rho0 = (weight_density_p(x0, q, p2c, Ng, N, dx) if deposit != 'window-det' else
            ops.pypic_weight(np.asarray(x0)[:N], np.asarray(q)[:N], None, p2c, Ng, dx, reproducible=True))
    phi0 = solve_poisson_p(dx, Ng, rho0, np.zeros(Ng))
    phi0 = phi0 - np.max(phi0)
    E0 = -differentiate_p(phi0, dx, Ng)
    if sort_every is None:
        sort_every = 8 if N >= (1 << 17) else 0
    sim = PeriodicImplicitSim(N, Ng, dx, dt, L, p2c, q=-e, m=me, tol=tol, maxiter=maxiter, sort_every=sort_every,
                              deposit=deposit)
    sim.upload(x0, v0, E0)
    mpl, plt = get_plt()
    for t in range(T):
        print('t: ', t)
        k, r = sim.push()
        print("Iterations: ", k)
        print("Residual  : ", r)
        d = sim.diagnostics(me)
        TT.append(t * dt)
        EE.append(d["EE"])
        KE.append(d["KE"])
        print("Total Energy: ", d["EE"] + d["KE"])
        j_bias.append(d["jbias"])
        if tracer < N:
            slot = sim.slot_of(tracer)
            trajectory_x.append(float(sim.x0[slot].item()) % L)        # the store wraps lazily
            trajectory_v.append(float(sim.v0[slot].item()) / np.sqrt(kBTe / me))
        if plt is not None and (t % nplot == 0):
            st = sim.download()
            fig = plt.figure(1); plt.clf()
            ax = fig.subplots(2, 2)
            ax[0, 0].hist2d(st["x0"], st["v0"] / np.sqrt(kBTe / me), bins=(100, 50), range=[[0.0, L], [-Vmax, Vmax]])
            ax[0, 1].hist(st["v0"] / np.sqrt(kBTe / me), bins=200, orientation='horizontal', density=True)
            ax[1, 1].semilogy(np.array(TT) * wp, EE, linewidth=lw)
            ax[1, 0].plot(X[:-1], st["E0"], linewidth=lw)
            plt.savefig(outdir + '/summary_' + str(t))
    sim.check()
    st = sim.download()
    np.savetxt(outdir + '/E2.txt', EE)
    np.savetxt(outdir + '/J.txt', st["j0"])
    with open(outdir + '/parameters.out', 'w+') as output_file:
        for name, val in (('wp', wp), ('Te', Te), ('G', growth_rate), ('tau', 1.0 / wp), ('p2c', p2c), ('dt', dt),
                          ('dx', dx), ('Ng', Ng), ('L', L + dx)):
            print(name, val, file=output_file)
    if result is not None:
        result.update(st, EE=np.array(EE), KE=np.array(KE), TT=np.array(TT), j_bias=np.array(j_bias))
    return EE


def explicit_pic(T, nplot):
    raise NotImplementedError("pypic.explicit_pic is dead code in the reference (NameError on Ng, dx, system; "
                              "wrong initialize_p arity, pypic.py:653-812); the explicit loop is PIC_L.main")


def main(T, nplot, N=1000000, Ng=200, dt=1e-5, density=1e5, perturbation=0.8, Kp=1, system='landau-damping',
         tol=1e-3, maxiter=20, outdir='plots', result=None, sort_every=None, deposit='warp'):
    """pypic.main (pypic.py:814-863); keyword arguments default to its hard-coded literals."""
    Ti = 0.1 * 11600.
    Te = 100.0 * 11600.
    L = 22.0 * np.sqrt(kb * Te * epsilon0 / e / e / density)
    Vmax = 8.0
    Nv = Ng // 2
    implicit_pic(T, nplot, system, density, perturbation, Kp, N, Ng, Nv, Vmax, dt, Ti, Te, L, tol, maxiter,
                 outdir=outdir, result=result, sort_every=sort_every, deposit=deposit)


if __name__ == '__main__':
    main(100, 10)
