"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of pyPIC's per-timestep
hot path.  Never imported by the product (pypic_b200/, drop-in modules, the GPU
arm of bench.py); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may use it, and only as the checker /
baseline.

Parity status: PINNED.  Every function here is checked in tests/test_oracle.py
against (a) the known-answer doctests the reference ships (pygcpic.py, SURVEY.md
section 4), (b) golden vectors produced by executing the reference's own code
through oracle/refshim.py (tests/golden/*.npz, generator oracle/make_golden.py)
and (c) live against /root/reference when it is mounted.

Each function cites the reference lines it follows.  Operation order is kept
exactly as written in the reference so that per-particle results are
bit-identical; scatter loops are restated with ``np.add.at`` on *interleaved*
(left,right) index streams, which NumPy executes unbuffered in array order, i.e.
in the same serial particle order as the reference's ``for i in range(N)``.

Third-party arithmetic that is not under /root/reference:
  * scipy.sparse.linalg.spsolve / inv (SuperLU; SciPy unpinned, 1.18.1 here):
    restated as a gauge-fixed direct tridiagonal solve; equal after the caller's
    ``phi - max(phi)`` to ~1e-12.
  * scipy.sparse.linalg.bicgstab (default rtol, x0=phi) inside the Newton loop of
    pygcpic.Grid.solve_for_phi_dirichlet_boltzmann: restated as an exact
    tridiagonal Newton step (same fixed point; documented phi tolerance).
  * numba fastmath kernels (numba unpinned, 0.65.0 here): restated without FMA
    contraction; agreement <= a few ulp, cell indices exact.
  * NumPy legacy MT19937 RandomState: used directly (stream frozen by NumPy).
"""
import numpy as np

epsilon0 = 8.854E-12
e = 1.602E-19
mp = 1.67E-27
me = 9.11E-31
kb = 1.38E-23


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _scatter_serial(out, idx_pairs, val_pairs):
    """Serial-order accumulate.  idx_pairs/val_pairs: (N,2) arrays holding the
    (left,right) targets of particle i; flattened row-major this is exactly the
    order of ``for i: out[iL]+=..; out[iR]+=..``."""
    np.add.at(out, idx_pairs.reshape(-1), val_pairs.reshape(-1))
    return out


def solve_tridiagonal(a, b, c, d):
    """Thomas algorithm, a=sub (a[0] unused), b=diag, c=super (c[-1] unused)."""
    n = len(d)
    cp = np.zeros(n)
    dp = np.zeros(n)
    cp[0] = c[0] / b[0]
    dp[0] = d[0] / b[0]
    for i in range(1, n):
        den = b[i] - a[i] * cp[i - 1]
        cp[i] = c[i] / den if i < n - 1 else 0.0
        dp[i] = (d[i] - a[i] * dp[i - 1]) / den
    x = np.zeros(n)
    x[-1] = dp[-1]
    for i in range(n - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return x


def solve_tridiagonal_fast(a, b, c, d):
    """Same system through LAPACK (scipy.linalg.solve_banded) for large n."""
    from scipy.linalg import solve_banded
    n = len(d)
    ab = np.zeros((3, n))
    ab[0, 1:] = c[:-1]
    ab[1, :] = b
    ab[2, :-1] = a[1:]
    return solve_banded((1, 1), ab, d)


# --------------------------------------------------------------------------
# pypic.py -- periodic implicit PIC (numba kernels)
# --------------------------------------------------------------------------
def pypic_indices_weights(x, Ng, dx, weight_by_division):
    """pypic.py:45-53 / 110-118 / 157-165.  iL=int(x*(1/dx)) (truncation),
    iR=int((x*(1/dx)+1)%Ng) (float modulo), wR=(x%dx)*idx (interpolate_p,
    weight_current_p) or (x%dx)/dx (weight_density_p)."""
    idx = (1. / dx)
    index_L = x * idx
    index_R = (index_L + 1) % Ng
    if weight_by_division:
        w_R = (x % dx) / dx
    else:
        w_R = (x % dx) * idx
    w_L = 1. - w_R
    return index_L.astype(np.int64), index_R.astype(np.int64), w_L, w_R


def pypic_interpolate_p(F, x, Ng, N, dx):
    """pypic.py:28-61."""
    iL, iR, wL, wR = pypic_indices_weights(x, Ng, dx, False)
    return F[iL] * wL + F[iR] * wR


def pypic_smooth_field_p(F):
    """pypic.py:63-76."""
    return (np.roll(F, -1) + 2.0 * F + np.roll(F, 1)) * 0.25


def pypic_weight_current_p(x, q, v, p2c, Ng, N, dx):
    """pypic.py:91-136.  p2c is truncated to int32 at the numba call boundary
    (SURVEY.md C11)."""
    p2c = int(p2c)
    iL, iR, wL, wR = pypic_indices_weights(x, Ng, dx, False)
    idx = (1. / dx)
    j_i = q * v * p2c * idx
    j = np.zeros(Ng)
    _scatter_serial(j, np.stack([iL, iR], 1), np.stack([j_i * wL, j_i * wR], 1))
    return j


def pypic_weight_density_p(x, q, p2c, Ng, N, dx):
    """pypic.py:138-183."""
    p2c = int(p2c)
    iL, iR, wL, wR = pypic_indices_weights(x, Ng, dx, True)
    idx = (1. / dx)
    q_i = q * p2c * idx
    rho = np.zeros(Ng)
    _scatter_serial(rho, np.stack([iL, iR], 1), np.stack([q_i * wL, q_i * wR], 1))
    return rho


def pypic_differentiate_p(F, dx, Ng):
    """pypic.py:185-214 (periodic centred difference, +dF/dx)."""
    idx_2 = (0.5 / dx)
    return (np.roll(F, -1) - np.roll(F, 1)) * idx_2


def pypic_solve_poisson_p(dx, Ng, rho, phi0=None):
    """pypic.py:359-382 + 337-357.  The periodic [1,-2,1] matrix is singular; the
    reference hands it to SuperLU and the caller subtracts max(phi)
    (pypic.py:553).  Restated gauge-fixed: phi[Ng-1]=0, solve the (Ng-1)
    Dirichlet system.  Compare only after ``phi - max(phi)``."""
    dx2 = dx * dx
    c0 = -np.average(rho) / epsilon0
    c2 = rho / epsilon0
    rhs = -dx2 * c0 - dx2 * c2
    n = Ng - 1
    a = np.ones(n)
    b = -2. * np.ones(n)
    c = np.ones(n)
    phi = np.zeros(Ng)
    phi[:n] = solve_tridiagonal_fast(a, b, c, rhs[:n])
    return phi


def pypic_particle_push_p(x0, v0, q, m, E0, j0, N, Ng, p2c, dx, dt, L, tol, maxiter):
    """pypic.py:216-300.  Returns (x1, v1, E1, j1, k, r)."""
    q_m = q / m
    Es = E0
    xs = x0
    r = 1.0
    k = 0
    x1 = x0
    v1 = v0
    E1 = E0
    j1 = j0
    while (r > tol) & (k < maxiter):
        E_interp = pypic_interpolate_p(pypic_smooth_field_p(Es), xs, Ng, N, dx)
        x1 = x0 + dt * v0 + dt * dt * (q_m) * E_interp * 0.5
        v1 = v0 + dt * (q_m) * E_interp
        xh = (x0 + x1) * 0.5
        vh = (v0 + v1) * 0.5
        xh = xh % L
        jh = pypic_weight_current_p(xh, q, vh, p2c, Ng, N, dx)
        x1 = x1 % L
        j1 = pypic_weight_current_p(x1, q, v1, p2c, Ng, N, dx)
        E1 = E0 + (dt / epsilon0) * (np.average(jh) - pypic_smooth_field_p(jh))
        Eh = (E1 + E0) * 0.5
        r = np.sum((Es - Eh) ** 2)
        Es = Eh
        xs = xh
        k += 1
    return x1, v1, E1, j1, k, r


# --------------------------------------------------------------------------
# PIC_L_DD.py -- bounded two-species implicit sheath
# --------------------------------------------------------------------------
def dd_index_weights(x, dx):
    """PIC_L_DD.py:33-36, 44-46: index=floor(x/dx) (true division),
    wR=(x%dx)/dx."""
    index = np.floor(x / dx)
    wR = (x % dx) / dx
    wL = 1. - wR
    return index.astype(np.int64), wL, wR


def dd_interpolateField(F, x, Ng, dx):
    """PIC_L_DD.py:32-39 (vectorised over x)."""
    index, wL, wR = dd_index_weights(np.asarray(x, dtype=np.float64), dx)
    return wL * F[index] + wR * F[index + 1]


def dd_weightCurrents(x, q, v, p2c, Ng, N, dx, dt, active):
    """PIC_L_DD.py:41-68 -- CIC current of active particles plus the wall-charge
    terms of absorbed ones (active==-1 left wall, active==0 right wall), then the
    edge fold j[0]+=j[1]; j[-1]+=j[-2].  Serial particle order."""
    j = np.zeros(Ng)
    idx = (1. / dx)
    act = (active == 1)
    left = (active == -1)
    right = (active == 0)
    with np.errstate(invalid="ignore"):
        index = np.floor(x / dx)
        wR = (x % dx) / dx
    wL = 1. - wR
    tL = np.zeros(N, dtype=np.int64)
    tR = np.zeros(N, dtype=np.int64)
    vL = np.zeros(N)
    vR = np.zeros(N)
    ia = index[act].astype(np.int64)
    tL[act] = ia
    tR[act] = ia + 1
    vL[act] = q[act] * v[act] * p2c * wL[act] * idx
    vR[act] = q[act] * v[act] * p2c * wR[act] * idx
    tL[left] = 0
    tR[left] = 0
    vL[left] = dx * q[left] * p2c / dt
    tL[right] = Ng - 1
    tR[right] = Ng - 1
    vL[right] = -dx * q[right] * p2c / dt
    _scatter_serial(j, np.stack([tL, tR], 1), np.stack([vL, vR], 1))
    j[0] += j[1]
    j[-1] += j[-2]
    return j


def dd_weightDensities(x, q, p2c, Ng, N, dx, active):
    """PIC_L_DD.py:70-88."""
    rho = np.zeros(Ng)
    idx = (1. / dx)
    act = (active == 1)
    xa = x[act]
    index, wL, wR = dd_index_weights(xa, dx)
    vL = q[act] * p2c * wL * idx
    vR = q[act] * p2c * wR * idx
    _scatter_serial(rho, np.stack([index, index + 1], 1), np.stack([vL, vR], 1))
    return rho


def dd_differentiateField(F, dx, Ng):
    """PIC_L_DD.py:192-203 (returns -dF/dx; one-sided at both ends)."""
    dF = np.zeros(Ng)
    dF[1:-1] = -(F[2:] - F[:-2]) / dx * 0.5
    dF[-1] = -(F[-1] - F[-2]) / dx
    dF[0] = -(F[1] - F[0]) / dx
    return dF


def dd_integrateField(F, dx, Ng):
    """PIC_L_DD.py:205-214: IF[i] = -trapz(F[:i+1], dx).  O(Ng) restatement of
    the reference's O(Ng^2) loop (summation order differs: tolerance, not bits)."""
    IF = np.zeros(Ng)
    seg = dx * (F[1:] + F[:-1]) / 2.0
    IF[1:] = -np.cumsum(seg)
    return IF


def dd_smoothField(F):
    """PIC_L_DD.py:216-221."""
    Fs = (np.roll(F, -1) + 2.0 * F + np.roll(F, 1)) / 4.0
    Fs[0] = F[0]
    Fs[-1] = F[-1]
    return Fs


def dd_initialize_beam(N, density, dx, Ng, Te, Ti, L, rng=np.random):
    """PIC_L_DD.initialize('beam', ...) with perturbation=0, PIC_L_DD.py:223-314.
    Draw order: uniform scalar (279), u e-/i+ (285-286), v (288-289), w (291-292),
    x0 uniform N (296)."""
    kBTe = kb * Te
    kBTi = kb * Ti
    h = N // 2
    m = np.zeros(N)
    q = np.zeros(N)
    species = np.zeros(N)
    m[:h] = np.ones(h) * me
    q[:h] = -np.ones(h) * e
    m[h:] = 1.0 * np.ones(h) * mp
    q[h:] = np.ones(h) * e
    species[:h] = 1
    species[h:] = 2
    rng.uniform(0.0, L)
    u0 = np.zeros(N)
    v0 = np.zeros(N)
    w0 = np.zeros(N)
    u0[:h] = rng.normal(0.0, np.sqrt(kBTe / m[:h]))
    u0[h:] = rng.normal(0.0, np.sqrt(kBTi / m[h:]))
    v0[:h] = rng.normal(0.0, np.sqrt(kBTe / m[:h]))
    v0[h:] = rng.normal(0.0, np.sqrt(kBTi / m[h:]))
    w0[:h] = rng.normal(0.0, np.sqrt(kBTe / m[:h]))
    w0[h:] = rng.normal(0.0, np.sqrt(kBTi / m[h:]))
    x0 = rng.uniform(0., L, N)
    return m, q, x0, u0, v0, w0, species, kBTe, kBTi


def dd_reinject(x0, u0, v0, w0, active, species, m, L, kBTe, kBTi, gamma, rng=np.random):
    """PIC_L_DD.py:419-450: thermostat loop (one uniform per ACTIVE particle even
    when gamma==0 because of the short-circuit ``and``) then re-initialisation
    of every inactive slot, in index order.  Mutates the arrays in place and
    returns the number of re-injected particles."""
    N = len(x0)
    if gamma == 0.0:
        n_act = int(np.count_nonzero(active == 1))
        if n_act:
            rng.uniform(0.0, 1.0, n_act)
    else:
        for i in range(N):
            if active[i] == 1 and rng.uniform(0.0, 1.0) < gamma:
                u0[i] = rng.normal(0.0, np.sqrt(kBTi / m[i]))
                v0[i] = rng.normal(0.0, np.sqrt(kBTi / m[i]))
                w0[i] = rng.normal(0.0, np.sqrt(kBTi / m[i]))
    dead = np.nonzero(active != 1)[0]
    for i in dead:
        kT = kBTi if species[i] == 2 else kBTe
        x0[i] = rng.uniform(0.0, L)
        u0[i] = rng.normal(0.0, np.sqrt(kT / m[i]))
        v0[i] = rng.normal(0.0, np.sqrt(kT / m[i]))
        w0[i] = rng.normal(0.0, np.sqrt(kT / m[i]))
        active[i] = 1
    return len(dead)


def dd_picard_step(x0, u0, v0, w0, q, m, active, E0, p2c, Ng, dx, dt, L, tol, maxiter,
                   t=0, vionout=None, trace=None):
    """One timestep of PIC_L_DD.main_i's Picard loop, PIC_L_DD.py:452-545.
    ``active`` is mutated (1 -> 0 right wall / -1 left wall).  Returns the
    committed (x1,u1,v1,w1,E1,j1) plus (k, r, phih)."""
    N = len(x0)
    Es = E0
    xs = x0
    r = 1.0
    k = 0
    qm = q / m
    x1 = np.zeros(N); u1 = np.zeros(N); v1 = np.zeros(N); w1 = np.zeros(N)
    E1 = E0
    j1 = np.zeros(Ng)
    phih = np.zeros(Ng)
    while (r > tol) & (k < maxiter):
        x1 = np.zeros(N); u1 = np.zeros(N); v1 = np.zeros(N); w1 = np.zeros(N)
        xh = np.zeros(N); uh = np.zeros(N)
        act = (active == 1)
        Ei = dd_interpolateField(Es, xs[act], Ng, dx)
        x1[act] = x0[act] + dt * u0[act] + dt * dt * qm[act] * Ei * 0.5
        u1[act] = u0[act] + dt * qm[act] * Ei
        v1[act] = v0[act]
        w1[act] = w0[act]
        xh[act] = (x0[act] + x1[act]) * 0.5
        uh[act] = (u0[act] + u1[act]) * 0.5
        # absorption, PIC_L_DD.py:494-505
        right = act & ((x0 >= L) | (xh >= L) | (x1 >= L))
        active[right] = 0
        left = (active == 1) & ((x0 <= 0.0) | (xh <= 0.0) | (x1 <= 0.0))
        active[left] = -1
        if vionout is not None and t > 2000:
            h = N // 2
            for i in np.nonzero((right | left)[:h])[0]:
                vionout.append(u0[i] if right[i] else -u0[i])
        jh = dd_weightCurrents(xh, q, uh, p2c, Ng, N, dx, dt, active)
        j1 = dd_weightCurrents(x1, q, u1, p2c, Ng, N, dx, dt, active)
        E1 = E0 + (dt / epsilon0) * (np.average(jh) - jh)
        Eh = (E1 + E0) * 0.5
        phih = dd_integrateField(Eh, dx, Ng)
        phih = phih - np.max(phih)
        r = np.linalg.norm(Es - Eh)
        if trace is not None:
            trace.append(dict(k=k, r=float(r), jh=jh.copy(), j1=j1.copy(), Eh=Eh.copy(),
                              n_right=int(np.count_nonzero(active == 0)),
                              n_left=int(np.count_nonzero(active == -1))))
        Es = Eh
        xs = xh
        k += 1
    return x1, u1, v1, w1, E1, j1, k, r, phih


def dd_main_i(T, N=40000, Ng=51, dt=1E-12, dx=0.00001, Ti=10.0 * 11600., Te=10.0 * 11600.,
              density=1E19, gamma=0.0, tol=1E-5, maxiter=20, rng=np.random, record=None):
    """Restatement of PIC_L_DD.main_i's time loop (PIC_L_DD.py:316-551) without
    plotting.  Returns a dict of final state and time series."""
    L = dx * (Ng - 1)
    p2c = (L) * density / N
    m, q, x0, u0, v0, w0, species, kBTe, kBTi = dd_initialize_beam(N, density, dx, Ng, Te, Ti, L, rng)
    active = np.ones(N)
    E0 = dd_differentiateField(np.zeros(Ng), dx, Ng)
    j0 = dd_weightCurrents(x0, q, u0, p2c, Ng, N, dx, dt, active)
    EE = []; KE = []; TT = []; jbias = []; vionout = []; iters = []; resid = []; ninj = []
    for t in range(T + 1):
        ninj.append(dd_reinject(x0, u0, v0, w0, active, species, m, L, kBTe, kBTi, gamma, rng))
        x1, u1, v1, w1, E1, j1, k, r, phih = dd_picard_step(
            x0, u0, v0, w0, q, m, active, E0, p2c, Ng, dx, dt, L, tol, maxiter, t, vionout)
        E0 = E1; x0 = x1; u0 = u1; v0 = v1; w0 = w1; j0 = j1
        iters.append(k); resid.append(r)
        EE.append(np.sum(epsilon0 * E0 * E0 * dx / 2.))
        KE.append(np.sum(me * u0 * u0 / 2.))
        TT.append(t * dt)
        jbias.append(np.average(j0))
        if record is not None:
            record(t, x0, u0, v0, w0, active, E0, j0, phih)
    return dict(x0=x0, u0=u0, v0=v0, w0=w0, active=active, E0=E0, j0=j0, phih=phih,
                EE=np.array(EE), KE=np.array(KE), TT=np.array(TT), jbias=np.array(jbias),
                vionout=np.array(vionout), iters=np.array(iters), resid=np.array(resid),
                ninj=np.array(ninj), p2c=p2c, L=L)


# --------------------------------------------------------------------------
# PIC_L.py -- periodic explicit leapfrog + Poisson every step
# --------------------------------------------------------------------------
def l_interpolateFieldPeriodic(F, x, Ng, dx):
    """PIC_L.py:39-46: index=int(floor(x/dx))%(Ng+1)."""
    index = (np.floor(x / dx).astype(np.int64)) % (Ng + 1)
    wR = (x % dx) / dx
    wL = 1. - wR
    return wL * F[index] + wR * F[index + 1]


def l_weightDensitiesPeriodic(x, q, p2c, Ng, N, dx):
    """PIC_L.py:100-118 (Ng+1 nodes; fold rho[-1]=rho[0]+rho[-1]; rho[0]=rho[-1])."""
    rho = np.zeros(Ng + 1)
    index = (np.floor(x / dx) % (Ng + 1)).astype(np.int64)
    wR = (x % dx) / dx
    wL = 1. - wR
    idx = (1. / dx)
    _scatter_serial(rho, np.stack([index, index + 1], 1),
                    np.stack([q * p2c * wL * idx, q * p2c * wR * idx], 1))
    rho[-1] = rho[0] + rho[-1]
    rho[0] = rho[-1]
    return rho


def l_weightCurrentsPeriodic(x, q, v, p2c, Ng, N, dx):
    """PIC_L.py:62-80 (fold j[0]=j[-1]+j[0]; j[-1]=j[0])."""
    j = np.zeros(Ng + 1)
    index = (np.floor(x / dx) % (Ng + 1)).astype(np.int64)
    wR = (x % dx) / dx
    wL = 1. - wR
    idx = (1. / dx)
    _scatter_serial(j, np.stack([index, index + 1], 1),
                    np.stack([q * v * p2c * wL * idx, q * v * p2c * wR * idx], 1))
    j[0] = j[-1] + j[0]
    j[-1] = j[0]
    return j


def l_solvePoissonPeriodicElectronsNeutralized(dx, Ng, rho):
    """PIC_L.py:208-220 with the (Ng+1)x(Ng+1) periodic matrix of 120-132.
    Singular system; gauge-fixed here (last unknown = 0); compare after -max."""
    n1 = Ng + 1
    dx2 = dx * dx
    c0 = -np.average(rho) / epsilon0
    c2 = rho / epsilon0
    rhs = -dx2 * c0 - dx2 * c2
    n = n1 - 1
    phi = np.zeros(n1)
    phi[:n] = solve_tridiagonal_fast(np.ones(n), -2. * np.ones(n), np.ones(n), rhs[:n])
    return phi


def l_differentiateFieldPeriodic(F, dx, Ng):
    """PIC_L.py:235-246 (Ng+1 nodes; returns -dF/dx)."""
    dF = np.zeros(Ng + 1)
    dF[1:-1] = -(F[2:] - F[:-2]) / dx * 0.5
    dF[-1] = -(F[0] - F[-2]) / dx * 0.5
    dF[0] = -(F[1] - F[-1]) / dx * 0.5
    return dF


def l_pushParticlesExplicit(x, v, q, m, N, Ng, dt, dx, E):
    """PIC_L.py:248-259 (kick-drift-kick, same E for both half kicks)."""
    E_interp = l_interpolateFieldPeriodic(E, x, Ng, dx)
    vhalf = v + (q / m) * (dt * 0.5) * E_interp
    xout = x + vhalf * dt
    vout = vhalf + (q / m) * (dt * 0.5) * E_interp
    return xout, vout


def l_explicit_step(x, v, q, m, p2c, Ng, N, dx, dt, L):
    """One pass of the PIC loop PIC_L.py:762-768.  Returns x,v,rho,phi,E."""
    rho = l_weightDensitiesPeriodic(x, q, p2c, Ng, N, dx)
    phi = l_solvePoissonPeriodicElectronsNeutralized(dx, Ng, rho)
    phi = phi - np.max(phi)
    E = l_differentiateFieldPeriodic(phi, dx, Ng)
    x, v = l_pushParticlesExplicit(x, v, q, m, N, Ng, dt, dx, E)
    x = x % (L + dx)
    return x, v, rho, phi, E


# --------------------------------------------------------------------------
# pygcpic.py -- Particle / Grid (SoA restatement; r is (N,7))
# --------------------------------------------------------------------------
def gc_gather_mirrored(E, x, dx):
    """pygcpic.py:344-347 -- NOTE the mirrored weights: the LEFT node gets the
    fractional distance w_l=(x%dx)/dx."""
    ind = np.floor(x / dx).astype(np.int64)
    w_l = (x % dx) / dx
    w_r = 1.0 - w_l
    return E[ind] * w_l + E[ind + 1] * w_r


def gc_push_6D(r, Ex, B, charge_state, m, dt):
    """pygcpic.py:460-507, Boris-Buneman 1D3V.  r: (N,7) modified copy returned."""
    r = r.copy()
    constant = 0.5 * dt * charge_state * 1.602e-19 / m
    r[:, 3] += constant * Ex
    tx = constant * B[0]
    ty = constant * B[1]
    tz = constant * B[2]
    t2 = tx * tx + ty * ty + tz * tz
    sx = 2. * tx / (1. + t2)
    sy = 2. * ty / (1. + t2)
    sz = 2. * tz / (1. + t2)
    vfx = r[:, 3] + r[:, 4] * tz - r[:, 5] * ty
    vfy = r[:, 4] + r[:, 5] * tx - r[:, 3] * tz
    vfz = r[:, 5] + r[:, 3] * ty - r[:, 4] * tx
    r[:, 3] += vfy * sz - vfz * sy
    r[:, 4] += vfz * sx - vfx * sz
    r[:, 5] += vfx * sy - vfy * sx
    r[:, 3] += constant * Ex
    r[:, 0] += r[:, 3] * dt
    r[:, 1] += r[:, 4] * dt
    r[:, 2] += r[:, 5] * dt
    r[:, 6] += dt
    return r


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1)


def gc_transform_6D_to_GC(r, B, charge_state, m):
    """pygcpic.py:509-551 (including the reference's ``*e`` in rl_hat)."""
    r = r.copy()
    x = r[:, 0:3].copy()
    v = r[:, 3:6].copy()
    B2 = B[0] ** 2 + B[1] ** 2 + B[2] ** 2
    b = B / np.sqrt(B2)
    # v.dot(b): BLAS ddot of 3 elements; restated as an ordered sum
    vpar_mag = v[:, 0] * b[0] + v[:, 1] * b[1] + v[:, 2] * b[2]
    vpar = vpar_mag[:, None] * b[None, :]
    wc = np.abs(charge_state) * e * np.sqrt(B2) / m
    vperp = v - vpar
    vperp_mag = np.sqrt(vperp[:, 0] ** 2 + vperp[:, 1] ** 2 + vperp[:, 2] ** 2)
    vperp_hat = vperp / vperp_mag[:, None]
    mu = 0.5 * m * vperp_mag ** 2 / np.sqrt(B2)
    rl_mag = vperp_mag / wc
    rl_hat = (-np.sign(charge_state) * e)[:, None] * _cross(vperp_hat, np.broadcast_to(b, vperp_hat.shape))
    rl = rl_mag[:, None] * rl_hat
    r[:, 0:3] = x - rl
    r[:, 3] = vpar_mag
    r[:, 4] = mu
    return r


def gc_transform_GC_to_6D(r, B, charge_state, m, a):
    """pygcpic.py:553-596.  ``a`` = the (N,3) uniform(0,1) draws of line 583.
    Keeps the reference's ``aperp = a - a.dot(b)`` (scalar subtracted)."""
    r = r.copy()
    X = r[:, 0:3].copy()
    vpar_mag = r[:, 3].copy()
    mu = r[:, 4].copy()
    B2 = B[0] ** 2 + B[1] ** 2 + B[2] ** 2
    b = B / np.sqrt(B2)
    vperp_mag = np.sqrt(2.0 * mu * np.sqrt(B2) / m)
    wc = np.abs(charge_state) * e * np.sqrt(B2) / m
    rl_mag = vperp_mag / wc
    adotb = a[:, 0] * b[0] + a[:, 1] * b[1] + a[:, 2] * b[2]
    aperp = a - adotb[:, None]
    aperp_mag = np.sqrt(aperp[:, 0] ** 2 + aperp[:, 1] ** 2 + aperp[:, 2] ** 2)
    bperp_hat = aperp / aperp_mag[:, None]
    rl = rl_mag[:, None] * bperp_hat
    x = X + rl
    vperp_hat = _cross(np.broadcast_to(b, bperp_hat.shape), bperp_hat)
    v = vpar_mag[:, None] * b[None, :] + vperp_mag[:, None] * vperp_hat
    r[:, 0:3] = x
    r[:, 3:6] = v
    return r


def gc_eom_GC(r, E, B, charge_state, m):
    """pygcpic.py:616-645.  E: (N,3) or (3,), B: (3,)."""
    B2 = B[0] ** 2 + B[1] ** 2 + B[2] ** 2
    b0 = B[0] / np.sqrt(B2)
    b1 = B[1] / np.sqrt(B2)
    b2 = B[2] / np.sqrt(B2)
    wc = np.abs(charge_state) * e * np.sqrt(B2) / m
    rho = r[:, 3] / wc
    E = np.broadcast_to(E, (r.shape[0], 3))
    rdot = np.zeros_like(r)
    rdot[:, 0] = (E[:, 1] * B[2] - E[:, 2] * B[1]) / B2 + r[:, 3] * b0
    rdot[:, 1] = (E[:, 2] * B[0] - E[:, 0] * B[2]) / B2 + r[:, 3] * b1
    rdot[:, 2] = (E[:, 0] * B[1] - E[:, 1] * B[0]) / B2 + r[:, 3] * b2
    rdot[:, 3] = (E[:, 0] * r[:, 0] + E[:, 1] * r[:, 1] + E[:, 2] * r[:, 2]) / np.sqrt(B2) / rho
    return rdot


def gc_push_GC(r, E, B, charge_state, m, dt):
    """pygcpic.py:598-614 classic RK4 on the 7-vector; E,B frozen over the step."""
    r0 = r
    k1 = dt * gc_eom_GC(r0, E, B, charge_state, m)
    k2 = dt * gc_eom_GC(r0 + k1 / 2., E, B, charge_state, m)
    k3 = dt * gc_eom_GC(r0 + k2 / 2., E, B, charge_state, m)
    k4 = dt * gc_eom_GC(r0 + k3, E, B, charge_state, m)
    out = r + (k1 + 2. * k2 + 2. * k3 + k4) / 6.
    out[:, 6] += dt
    return out


def gc_apply_BCs_dirichlet(x, active, at_wall, length):
    """pygcpic.py:668-689 (strict inequalities)."""
    hit = (x < 0.0) | (x > length)
    active = np.where(hit, 0, active)
    at_wall = np.where(hit, 1, at_wall)
    return active, at_wall


def gc_weight_particles(x, charge_state, p2c, active, ng, dx):
    """pygcpic.py:868-883 -- rho and n deposit of active particles (serial)."""
    rho = np.zeros(ng)
    n = np.zeros(ng)
    act = (active == 1)
    xa = x[act]
    index_l = np.floor(xa / dx).astype(np.int64)
    w_r = (xa % dx) / dx
    w_l = 1.0 - w_r
    cs = charge_state[act]
    pc = p2c[act]
    _scatter_serial(rho, np.stack([index_l, index_l + 1], 1),
                    np.stack([cs * e * pc / dx * w_l, cs * e * pc / dx * w_r], 1))
    _scatter_serial(n, np.stack([index_l, index_l + 1], 1),
                    np.stack([pc / dx * w_l, pc / dx * w_r], 1))
    return rho, n


def gc_boltzmann_n0_update(phi, domain, Te, n, n0, p_old, added_particles, dt, ve):
    """pygcpic.py:889-904.  Returns (n0, rho0, p_old)."""
    trapz = getattr(np, "trapezoid", None) or np.trapz
    eta = np.exp(phi / Te / 11600.)
    if n0 is None:
        p_old = trapz(eta, domain)
        n0 = 0.9 * np.average(n)
    else:
        p_new = trapz(eta, domain)
        q_new = eta[0] + eta[-1]
        r_new = 2. * added_particles / dt
        fn = np.sqrt(ve * q_new * dt / p_new)
        n0 = n0 * ((1.0 - fn) * p_old / p_new + fn - fn * fn / 4.) + r_new * dt / p_new
        p_old = p_new
    return n0, n0 * e, p_old


def gc_smooth_rho(rho):
    """pygcpic.py:1055-1060."""
    s = (np.roll(rho, -1) + 2.0 * rho + np.roll(rho, 1)) * 0.25
    s[0] = rho[0]
    s[-1] = rho[-1]
    return s


def gc_differentiate_phi_to_E(phi, dx):
    """pygcpic.py:932-936."""
    E = np.zeros_like(phi)
    E[1:-1] = -(phi[2:] - phi[:-2]) / dx / 2.
    E[0] = -(phi[1] - phi[0]) / dx
    E[-1] = -(phi[-1] - phi[-2]) / dx
    return E


def gc_solve_for_phi_dirichlet(rho, dx):
    """pygcpic.py:987-1003: phi = -inv(A).rho*dx^2, rows 0/-1 identity; - min."""
    ng = len(rho)
    a = np.ones(ng); b = -2. * np.ones(ng); c = np.ones(ng)
    b[0] = 1.; c[0] = 0.; b[-1] = 1.; a[-1] = 0.
    phi = -solve_tridiagonal_fast(a, b, c, rho) * (dx * dx)
    return phi - np.min(phi)


def gc_solve_for_phi_dirichlet_boltzmann(rho, n0, Te, dx, tolerance=1e-9, iter_max=1000):
    """pygcpic.py:1005-1053 with the bicgstab step replaced by an exact
    tridiagonal solve (same Newton fixed point).  Returns (phi, iterations)."""
    ng = len(rho)
    phi = np.zeros(ng)
    dx2 = dx * dx
    c0 = e * n0 / epsilon0
    c1 = e / kb / Te
    c2 = rho / epsilon0
    residual = 1.0
    it = 0
    a = np.ones(ng); c = np.ones(ng)
    c[0] = 0.; a[-1] = 0.
    while (residual > tolerance) and (it < iter_max):
        Aphi = np.zeros(ng)
        Aphi[1:-1] = phi[:-2] - 2. * phi[1:-1] + phi[2:]
        Aphi[0] = phi[0]; Aphi[-1] = phi[-1]
        F = Aphi - dx2 * c0 * np.exp(c1 * phi) + dx2 * c2
        F[0] = 0.; F[-1] = 0.
        D = -dx2 * c0 * c1 * np.exp(c1 * phi)
        D[0] = -dx2 * c0 * c1
        D[-1] = -dx2 * c0 * c1
        b = -2. * np.ones(ng); b[0] = 1.; b[-1] = 1.
        dphi = solve_tridiagonal_fast(a, b + D, c, F)
        phi = phi - dphi
        residual = dphi.dot(dphi)
        it += 1
    return phi - np.min(phi), it


def gc_solve_for_phi_dirichlet_neumann_boltzmann(phi_start, n, n0, Te, dx, tolerance=1e-3, iter_max=100):
    """pygcpic.py:1062-1109.  Last row of A is [1,-4,3] (973-977); folded into
    tridiagonal form by eliminating the (ng-3) entry with row ng-2."""
    ng = len(n)
    phi = phi_start.copy()
    dx2 = dx * dx
    c0 = e * n0 / epsilon0
    c1 = e / kb / Te
    c2 = e * n / epsilon0
    residual = 1.0
    it = 0
    while (residual > tolerance) and (it < iter_max):
        Aphi = np.zeros(ng)
        Aphi[1:-1] = phi[:-2] - 2. * phi[1:-1] + phi[2:]
        Aphi[0] = phi[0]
        Aphi[-1] = 3. * phi[-1] - 4. * phi[-2] + 1. * phi[-3]
        F = Aphi - dx2 * c0 * np.exp(c1 * phi) + dx2 * c2
        F[0] = phi[0]
        F[-1] = 0.
        D = -dx2 * c0 * c1 * np.exp(c1 * phi)
        D[0] = -dx2 * c0 * c1
        D[-1] = 0.
        a = np.ones(ng); b = -2. * np.ones(ng) + D; c = np.ones(ng)
        b[0] = 1. + D[0]; c[0] = 0.
        # last row [1,-4,3] minus row ng-2 ([1,b[-2],1]) -> [0,-4-b[-2],2]
        rhs = F.copy()
        a[-1] = -4. - b[-2]
        b[-1] = 3. - 1.
        rhs[-1] = F[-1] - F[-2]
        dphi = solve_tridiagonal_fast(a, b, c, rhs)
        phi = phi - dphi
        residual = np.sqrt(dphi.dot(dphi))
        it += 1
    return phi - np.min(phi), it


def gc_particle_loop_decisions(active_entry, active_after, src_entry, src_after, source_N):
    """Exact restatement of the order-dependent rule of pygcpic.py:1498-1549.

    active_entry[i]  : particle i active at loop entry (bool)
    active_after[i]  : for initially-active i, whether it is still active after its
                       push + BC + mid-domain-exit logic
    src_entry[i]     : (Z==source and charge_state>0) at loop entry
    src_after[i]     : same predicate after the particle's own processing
                       (ionisation may change charge_state)
    The count at the moment particle i is visited =
        sum_{j<i} contributes_after[j] + sum_{j>=i} contributes_entry[j]
    where reactivated particles contribute 1 afterwards.  Sequential by nature;
    restated as a loop (oracle sizes only)."""
    N = len(active_entry)
    contrib = (active_entry & src_entry).astype(np.int64)
    count = int(contrib.sum())
    react = np.zeros(N, dtype=bool)
    dele = np.zeros(N, dtype=bool)
    for i in range(N):
        if active_entry[i]:
            new = 1 if (active_after[i] and src_after[i]) else 0
            count += new - contrib[i]
        else:
            if count < source_N:
                react[i] = True
                count += 1
            else:
                dele[i] = True
    return react, dele


def gc_compact_stable(arrays, delete_mask):
    """pygcpic.py:1552-1563: order-preserving removal of flagged indices."""
    keep = ~delete_mask
    return [a[keep] for a in arrays]


# --------------------------------------------------------------------------
# Monte-Carlo ionisation inside the particle loop (pygcpic.py:350-458, 1496-1549)
# --------------------------------------------------------------------------
_ION_TABLES = {     # (Z, charge_state) -> (Te [eV], R [cm^3/s]); data of pygcpic.py:373-383, 409-438
    (1, 0): ([8.626E-01, 1.011E+00, 2.178E+00, 3.539E+00, 5.146E+00, 7.069E+00, 9.410E+00, 1.231E+01, 1.598E+01,
              2.076E+01, 2.720E+01, 3.625E+01, 4.973E+01, 7.133E+01, 1.099E+02, 1.904E+02, 4.079E+02, 1.355E+03,
              1.390E+04, 8.595E+04],
             [7.553E-16, 8.291E-15, 1.714E-11, 2.470E-10, 9.985E-10, 2.398E-09, 4.412E-09, 6.940E-09, 9.869E-09,
              1.309E-08, 1.649E-08, 1.996E-08, 2.329E-08, 2.624E-08, 2.834E-08, 2.881E-08, 2.627E-08, 1.926E-08,
              8.109E-09, 3.829E-09]),
    (5, 0): ([8.626E-01, 1.329E+00, 2.160E+00, 3.140E+00, 4.314E+00, 5.741E+00, 7.508E+00, 9.746E+00, 1.267E+01,
              1.660E+01, 2.212E+01, 3.034E+01, 4.353E+01, 6.704E+01, 1.162E+02, 2.490E+02, 8.265E+02, 8.481E+03,
              8.669E+04],
             [1.057E-12, 3.996E-11, 5.912E-10, 2.458E-09, 6.083E-09, 1.155E-08, 1.878E-08, 2.767E-08, 3.806E-08,
              4.979E-08, 6.257E-08, 7.590E-08, 8.901E-08, 1.005E-07, 1.080E-07, 1.079E-07, 9.470E-08, 5.161E-08,
              2.159E-08]),
    (5, 1): ([8.612E-01, 1.869E+00, 4.028E+00, 6.547E+00, 9.522E+00, 1.308E+01, 1.741E+01, 2.276E+01, 2.956E+01,
              3.840E+01, 5.031E+01, 6.707E+01, 9.203E+01, 1.319E+02, 2.033E+02, 3.522E+02, 7.547E+02, 2.505E+03,
              2.571E+04, 8.582E+04],
             [1.375E-21, 1.396E-14, 2.693E-11, 3.643E-10, 1.393E-09, 3.188E-09, 5.629E-09, 8.554E-09, 1.182E-08,
              1.533E-08, 1.900E-08, 2.273E-08, 2.639E-08, 2.972E-08, 3.221E-08, 3.300E-08, 3.032E-08, 2.252E-08,
              9.306E-09, 5.538E-09]),
    (5, 2): ([1.366E+00, 2.819E+00, 6.073E+00, 9.875E+00, 1.436E+01, 1.972E+01, 2.624E+01, 3.432E+01, 4.456E+01,
              5.790E+01, 7.587E+01, 1.012E+02, 1.387E+02, 1.990E+02, 3.064E+02, 5.311E+02, 1.138E+03, 3.778E+03,
              3.877E+04, 8.602E+04],
             [1.230E-21, 2.871E-15, 5.524E-12, 7.439E-11, 2.824E-10, 6.401E-10, 1.117E-09, 1.677E-09, 2.293E-09,
              2.946E-09, 3.629E-09, 4.337E-09, 5.055E-09, 5.759E-09, 6.382E-09, 6.779E-09, 6.575E-09, 5.269E-09,
              2.483E-09, 1.829E-09]),
}


def gc_ionization_probability(x, p2c, Z, charge_state, n, dx, dt, temperature):
    """pygcpic.py:385-391 / 440-446 for ONE particle: np.interp'd rate, CIC-gathered density."""
    Te, R = _ION_TABLES[(int(Z), int(charge_state))]
    rate = np.interp(temperature, [T * 11600. for T in Te], [r_ / 1e6 for r_ in R])
    il = int(np.floor(x / dx))
    w_r = (x % dx) / dx
    w_l = 1.0 - w_r
    density = w_l * n[il] + w_r * n[il + 1]
    return density ** 2 * rate * dx * dt / p2c


def gc_ionizing_particle_loop(r, cs, m, p2c, Z, active, at_wall, from_wall, E, n, B, dt, dx, length, Te, source_Z,
                              source_N, source, src_p2c, src_m, time, rng=np.random):
    """One pass of the particle loop of pic_bca_aps (pygcpic.py:1496-1549) over SoA arrays,
    sequential like the reference (test sizes only).  Mutates the arrays; returns
    (hits, reactivated, deleted indices, ionised H, ionised B, mid exits, added p2c list)."""
    N = len(cs)
    nh = nr = nih = nib = nex = 0
    deleted, added = [], []
    for i in range(N):
        if active[i] == 1:
            Ex = gc_gather_mirrored(E, r[i:i + 1, 0], dx)
            r[i] = gc_push_6D(r[i:i + 1], Ex, B, cs[i], m[i], dt)[0]
            if r[i, 0] < 0.0 or r[i, 0] > length:
                active[i] = 0; at_wall[i] = 1
            tried = False
            if Z[i] == 1 and cs[i] == 0 and active[i] == 1:
                pr = gc_ionization_probability(r[i, 0], p2c[i], 1, 0, n, dx, dt, Te)
                if rng.uniform(0., 1.) < pr and cs[i] == 0.:
                    cs[i] = 1; added.append(p2c[i]); nih += 1
                tried = True
            if Z[i] == 5 and cs[i] < 3 and active[i] == 1:
                pr = gc_ionization_probability(r[i, 0], p2c[i], 5, cs[i], n, dx, dt, Te)
                if rng.uniform(0., 1.) < pr and cs[i] == 0.:
                    cs[i] += 1; added.append(p2c[i]); nib += 1
            if active[i] != 1 and at_wall[i]:
                nh += 1
            dxm = length / 8
            if from_wall[i] and (length / 2 - dxm < r[i, 0] < length / 2 + dxm):
                if active[i] == 1:
                    nex += 1
                active[i] = 0
        else:
            count = int(np.sum((Z == source_Z) & (active == 1) & (cs > 0)))
            if count < source_N:
                rn = next(source)
                r[i] = rn; r[i, 6] = time
                p2c[i] = src_p2c; m[i] = src_m; cs[i] = 1; Z[i] = source_Z
                active[i] = 1; at_wall[i] = 0; from_wall[i] = 0
                added.append(src_p2c); nr += 1
            else:
                deleted.append(i)
    return nh, nr, deleted, nih, nib, nex, added


# --------------------------------------------------------------------------
# Device-mode random draws: the counter-based Philox4x32-10 generator (Salmon et
# al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 -- a published algorithm, not
# reference code: the reference only has the MT19937 host draws) and the mappings the kernels of
# pypic_b200/csrc/init_kernels.cu and dd_kernels.cu apply to its output.  Pinned by Random123's
# known-answer vectors in tests/test_oracle.py.
# --------------------------------------------------------------------------
def philox4x32_10(c, k0, k1):
    """c: (n,4) uint32 counters; k0, k1: uint32 key words (scalars).  Returns the (n,4) uint32 output."""
    c = np.array(c, dtype=np.uint64).reshape(-1, 4)
    c0, c1, c2, c3 = (c[:, j].copy() for j in range(4))
    k0 = np.uint64(int(k0) & 0xffffffff); k1 = np.uint64(int(k1) & 0xffffffff)
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xffffffff)
    for _ in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask; k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack([c0, c1, c2, c3], 1).astype(np.uint32)


def u53(a, b):
    """uniform in (0,1) from two uint32 words (53 random bits), common.cuh:u53"""
    v = (a.astype(np.uint64) << np.uint64(21)) ^ (b.astype(np.uint64) >> np.uint64(11))
    return (v.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def u52(a, b):
    """dd_kernels.cu:u52 (the re-injection draws)"""
    v = (a.astype(np.uint64) << np.uint64(20)) ^ (b.astype(np.uint64) >> np.uint64(12))
    return (v.astype(np.float64) + 0.5) * (1.0 / 4503599627370496.0)


def dev_init_uniform_maxwellian(N, n_split, xlo, xhi, sigma, mean, seed, stream_id, goff=0):
    """pic_dev_init_uniform_maxwellian: x ~ U(xlo,xhi), v0 ~ N(mean_s, sigma_s), v1, v2 ~ N(0, sigma_s),
    Philox keyed by (seed, stream_id, global index).  Returns x, v0, v1, v2."""
    g = np.arange(N, dtype=np.uint64) + np.uint64(goff)
    lo, hi = (g & np.uint64(0xffffffff)), (g >> np.uint64(32))
    sid = int(stream_id)
    k0, k1 = int(seed) & 0xffffffff, ((int(seed) >> 32) & 0xffffffff) ^ 0x5bd1e995
    c = philox4x32_10(np.stack([lo, hi, np.full(N, sid & 0xffffffff, np.uint64), np.full(N, sid >> 32, np.uint64)], 1), k0, k1)
    d = philox4x32_10(np.stack([lo, hi, np.full(N, (sid & 0xffffffff) ^ 0x9e3779b9, np.uint64), np.full(N, sid >> 32, np.uint64)], 1),
                      k0, k1)
    sp = np.arange(N) >= n_split
    sig = np.where(sp, sigma[1], sigma[0]); mu = np.where(sp, mean[1], mean[0])
    x = xlo + u53(c[:, 0], c[:, 1]) * (xhi - xlo)

    def bm(u1, u2):
        r = np.sqrt(-2.0 * np.log(u1))
        return r * np.cos(np.pi * (2.0 * u2)), r * np.sin(np.pi * (2.0 * u2))
    z0, z1 = bm(u53(c[:, 2], c[:, 3]), u53(d[:, 0], d[:, 1]))
    z2, _ = bm(u53(d[:, 2], d[:, 3]), u53(c[:, 1], d[:, 2]))
    return x, mu + sig * z0, sig * z1, sig * z2


def dev_reinject_philox(gid, step, seed, L, sig):
    """dd_reinject_one (dd_kernels.cu): the device-mode re-injection draws of PIC_L_DD.py:429-450 for
    the particles with global keys gid at time step `step`.  Returns x, u, v, w."""
    g = np.asarray(gid, dtype=np.uint64); n = len(g)
    lo, hi = (g & np.uint64(0xffffffff)), (g >> np.uint64(32))
    s_lo, s_hi = int(step) & 0xffffffff, (int(step) >> 32) & 0xffffffff
    k0, k1 = int(seed) & 0xffffffff, (int(seed) >> 32) & 0xffffffff
    mk = lambda x3: philox4x32_10(np.stack([lo, hi, np.full(n, s_lo, np.uint64), np.full(n, s_hi ^ x3, np.uint64)], 1), k0, k1)
    c, d, gg = mk(0), mk(0x80000000), mk(0x40000000)
    twopi = 6.283185307179586
    r1 = np.sqrt(-2.0 * np.log(u52(c[:, 2], c[:, 3]))); t1 = twopi * u52(d[:, 0], d[:, 1])
    r2 = np.sqrt(-2.0 * np.log(u52(d[:, 2], d[:, 3]))); t2 = twopi * u52(gg[:, 0], gg[:, 1])
    return u52(c[:, 0], c[:, 1]) * L, sig * r1 * np.cos(t1), sig * r1 * np.sin(t1), sig * r2 * np.cos(t2)
