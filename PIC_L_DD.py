"""Drop-in replacement for the reference's PIC_L_DD.py (bounded two-species implicit
sheath): same module-level names, signatures and outputs, backed by the sm_100a CUDA
kernels of libpic_b200.so through pypic_b200.  `run_pypic_dd.py` of the reference drives
this module unchanged (tools/drive.py).  No CPU fallback: every numerical function runs
on the GPU; only configuration, the legacy-RNG draw service and I/O stay on the host.
"""
from __future__ import print_function

import os
import sys
import time

import numpy as np

from pypic_b200 import ops
from pypic_b200.plotting import get_plt
from pypic_b200.sheath import SheathSim

lw = 3.0

# physical constants (PIC_L_DD.py:26-30)
epsilon0 = 8.854E-12
e = 1.602E-19
mp = 1.67E-27
me = 9.11E-31
kb = 1.38E-23


def interpolateField(F, x, Ng, dx):
    """PIC_L_DD.py:32-39.  Accepts a scalar (like the reference) or an array of positions."""
    out = ops.dd_interpolate(F, x, Ng, dx)
    return float(out[0]) if np.ndim(x) == 0 else out


def weightCurrents(x, q, v, p2c, Ng, N, dx, dt, active):
    """PIC_L_DD.py:41-68."""
    return ops.dd_weight(x[:N], q[:N], v[:N], p2c, Ng, dx, dt, active[:N])


def weightDensities(x, q, p2c, Ng, N, dx, active):
    """PIC_L_DD.py:70-88."""
    return ops.dd_weight(x[:N], q[:N], None, p2c, Ng, dx, 1.0, active[:N])


def laplacian1DPeriodic(Ng):
    """PIC_L_DD.py:90-102 (matrix constructor; host-side helper)."""
    A = np.diag(np.ones(Ng - 1), -1) + np.diag(-2. * np.ones(Ng), 0) + np.diag(np.ones(Ng - 1), 1)
    A[0, -1] = 1.
    A[-1, 0] = 1.
    return A


def laplacian1D(Ng):
    """PIC_L_DD.py:104-114."""
    A = np.diag(np.ones(Ng - 1), -1) + np.diag(-2. * np.ones(Ng), 0) + np.diag(np.ones(Ng - 1), 1)
    A[0, 0] = 1.
    A[0, 1] = 0.
    A[0, 2] = 0.
    A[-1, -1] = -2.
    A[-1, -2] = 1.
    A[-1, -3] = 1.
    return A


def solvePoisson(dx, Ng, rho, kBT, tol, maxiter, phi0):
    """PIC_L_DD.py:116-147: Boltzmann-Newton solve on Ng nodes as written (reference node Ng/2, the
    three-entry last row of laplacian1D, `while resid > tol and k <= maxiter`); the whole Newton
    loop is one kernel launch, a PCR solve per iteration in place of scipy.sparse.linalg.inv."""
    return ops.newton_boltzmann_l(np.asarray(rho)[:Ng], phi0, dx, kBT, tol, maxiter, periodic=False)


def solvePoissonPeriodic(dx, Ng, rho, kBT, tol, maxiter, phi0):
    """PIC_L_DD.py:149-176: the periodic variant on Ng nodes (cyclic system by Sherman-Morrison)."""
    return ops.newton_boltzmann_l(np.asarray(rho)[:Ng], phi0, dx, kBT, tol, maxiter, periodic=True)


def solvePoissonPeriodicElectronsNeutralized(dx, Ng, rho, kBT, tol, maxiter, phi0):
    """PIC_L_DD.py:178-190.  The periodic matrix is singular; the gauge is phi[-1]=0 (callers
    subtract max(phi))."""
    return ops.poisson_periodic(rho, dx, subtract_max=False)


def differentiateField(F, dx, Ng):
    """PIC_L_DD.py:192-203."""
    return ops.differentiate(F, dx, 1)


def integrateField(F, dx, Ng):
    """PIC_L_DD.py:205-214 (prefix scan instead of the O(Ng^2) loop)."""
    return ops.integrate_field(F, dx)


def smoothField(F):
    """PIC_L_DD.py:216-221."""
    return ops.smooth(F, 1)


def initialize(system, N, density, Kp, perturbation, dx, Ng, Te, Ti, L, X):
    """PIC_L_DD.py:223-314 for system == 'beam' (the only one main_i uses).  Host-side: the
    draws come from the global legacy np.random stream in the reference's order."""
    wp = np.sqrt(e**2 * density / epsilon0 / me)
    K = Kp * np.pi / (L)
    kBTe = kb * Te
    kBTi = kb * Ti
    h = N // 2
    m = np.zeros(N); q = np.zeros(N); species = np.zeros(N)
    m[:h] = me; q[:h] = -e; species[:h] = 1
    m[h:] = 1.0 * mp; q[h:] = e; species[h:] = 2
    if system != 'beam':
        raise NotImplementedError("only the 'beam' initialiser is used by main_i (PIC_L_DD.py:341)")
    growth_rate = -np.sqrt(np.pi) * wp**4 / K**3 / np.sqrt(kBTe / me)**3 * \
        np.exp(- wp**2 / K**2 / np.sqrt(kBTe / me)**2 * np.exp(-3. / 2.))
    print('Growth rate: ', growth_rate)
    np.random.uniform(0.0, L)                      # discarded scalar draw, PIC_L_DD.py:279
    u0 = np.zeros(N); v0 = np.zeros(N); w0 = np.zeros(N)
    for arr in (u0, v0, w0):
        arr[:h] = np.random.normal(0.0, np.sqrt(kBTe / m[:h]))
        arr[h:] = np.random.normal(0.0, np.sqrt(kBTi / m[h:]))
    x0 = np.random.uniform(0., L, N)
    F = -np.cos(Kp * np.pi * X / (L)) + 1.0
    F = (N * perturbation) * F / np.sum(F)
    j = h - int(N * perturbation / 2)
    for i in range(Ng):
        for k in range(int(F[i])):
            x0[j] = np.random.uniform(X[i], X[i + 1])
            j += 1
    return m, q, x0, u0, v0, w0, species, kBTe, kBTi, growth_rate


def main_i(T, nplot, N=40000, Ng=51, dt=1E-12, dx=0.00001, Ti=10.0 * 11600., Te=10.0 * 11600., density=1E19,
           gamma=0.0, tol=1E-5, maxiter=20, outdir='.', rng='host', result=None, sort_every=None, vion_after=2000,
           deposit='window'):
    """PIC_L_DD.main_i (PIC_L_DD.py:316-644).  The positional signature is the reference's;
    the keyword arguments default to its hard-coded literals.  `result` (a dict) receives
    the time series and the final state.

    The particle store is re-sorted by cell every `sort_every` steps so that the loop runs on the
    fused TMA kernel; the sort carries every particle's original index, so the legacy-RNG
    re-injection draws (made in index order, :429-450), vionout (:497-503) and every array handed
    back are in the reference's particle numbering.  The thermostat's uniforms of a gamma == 0 run
    are skipped by an MT19937 jump-ahead instead of being generated (pypic_b200/rng.py).
    deposit='window-det' runs the REPRODUCIBLE build: fixed-point accumulation of the currents, stable radix sort
    with the original-index payload and fixed-order diagnostics sums -- two runs give bit-identical output."""
    perturbation = 0.0
    Kp = 1.0
    L = dx * (Ng - 1)
    X = np.linspace(0.0, L, Ng)
    wp = np.sqrt(e**2 * density / epsilon0 / me)
    invwp = 1. / wp
    K = Kp * np.pi / (L)
    p2c = (L) * density / N
    m, q, x0, u0, v0, w0, species, kBTe, kBTi, growth_rate = initialize('beam', N, density, Kp, perturbation, dx, Ng,
                                                                        Te, Ti, L, X)
    print("wp : ", wp, "[1/s]")
    print("dt : ", dt / invwp, " [w * tau]")
    print("tau: ", invwp, "[s]")
    print("k  : ", K, "[1/m]")
    print("p2c :", p2c)
    print("floating potential: ", (kb * Te / e) * (0.5) * np.log(mp / 2.0 / np.pi / me))

    sim = SheathSim(N, Ng, dx, dt, p2c, q=(-e, e), m=(me, mp), tol=tol, maxiter=maxiter, kBT=(kBTe, kBTi),
                    gamma=gamma, carry_vw=True, rng=rng, sort_every=sort_every, vion_after=vion_after, deposit=deposit)
    sim.upload(x0, u0, v0, w0)          # E0 = -d(phi0)/dx with phi0 == 0 (PIC_L_DD.py:386-388)
    mpl, plt = get_plt()
    KE, EE, TT, jbias = [], [], [], []
    # np.std(u0) printed at the top of a step (:417) and KE summed at the end of the previous one (:549) are
    # moments of the same velocities: the first Picard iteration of the step accumulates them while it streams
    # u0 (SheathSim.fused_moments), so they arrive with the step's one device->host read and 'kBTe' is printed
    # right after the step has run -- in the reference's output order (nothing else prints in between)
    # (device-mode re-injection and the reproducible build: a pass of their own before the step)
    sim.fused_moments = (rng == 'host')
    t_loop = time.perf_counter()
    with sim.draws.hold():                           # the legacy stream's state stays in C for the duration of the loop
        for t in range(T + 1):
            print('t: ', t)
            pre = None if sim.fused_moments else sim.moments()
            k, r = sim.step()
            m1, m2 = sim.pre_step_moments() if pre is None else pre
            print('kBTe: ', sim.kBTe_from(m1, m2))
            print("Iterations: ", k)
            print("r: ", r)
            if t > 0:
                KE.append(me / 2. * m2)              # KE of step t-1 (:549)
            d = sim.step_stats()
            EE.append(d["EE"]); jbias.append(d["jbias"]); TT.append(t * dt)
            if plt is not None and (t % nplot == 0):
                st = sim.download()
                h = N // 2
                for fig, sl, name, size in ((1, slice(h, N), 'ps_i_', 2.0), (4, slice(0, h), 'ps_e_', 0.5)):
                    plt.figure(fig); plt.clf()
                    uu = st["u0"][sl]
                    plt.scatter(st["x0"][sl], np.sign(uu) * uu * uu * 0.5 * m[sl] / e, s=size)
                    plt.axis([0.0, L, -100.0, 100.0])
                    plt.savefig('plots/' + name + str(t))
                plt.figure(3); plt.clf()
                plt.plot(X, st["E0"], linewidth=lw)
                plt.savefig('plots/e_' + str(t))
    KE.append(me / 2. * sim.moments()[1])            # KE of the last step
    sim.check()
    t_loop = time.perf_counter() - t_loop
    if os.environ.get("PIC_TIMING"):
        sys.stderr.write("PIC_L_DD.main_i: time loop %d steps, %d particles: %.4f s = %.4e particle-steps/s (%.3f ms/step, "
                         "%d sorts, MT jumps %d, prefetched %d)\n" % (T + 1, N, t_loop, N * (T + 1) / t_loop,
                                                                      1e3 * t_loop / (T + 1), sim._sorts, sim.draws.jumps,
                                                                      sim.draws.prefetch_hits))
    st = sim.download()
    np.savetxt(outdir + '/vionout.txt', sim.collect_vionout())
    np.savetxt(outdir + '/E0.txt', st["E0"])
    np.savetxt(outdir + '/jb.txt', jbias)
    if result is not None:
        result.update(st, EE=np.array(EE), KE=np.array(KE), TT=np.array(TT), jbias=np.array(jbias),
                      phih=sim.phi(), p2c=p2c, loop_seconds=t_loop)
# end main_i


if __name__ == '__main__':
    main_i(1000, 10)
