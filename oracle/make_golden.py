"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING THE
REFERENCE'S OWN CODE (through oracle/refshim.py) on seeded inputs.

Run in the authoring container (needs /root/reference):
    python -m oracle.make_golden
The committed .npz files are what travels to the GPU box; this script is the
recipe that made them.  The only edits applied to reference sources are the
import shims documented in oracle/refshim.py plus, for the whole-loop runs,
LITERAL OVERRIDES of the hard-coded problem size inside main_i()/main()
(e.g. ``N = 40000`` -> ``N = 2000``) so the pure-Python loops finish quickly;
the arithmetic is untouched.
"""
import contextlib
import io
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def adversarial_positions(Ng, dx, L, rng):
    """Random positions + node-aligned positions (x = k*dx) and
    their floating-point neighbours, all inside [0, L)."""
    k = np.arange(0, Ng - 1)
    nodes = k * dx
    near = np.concatenate([np.nextafter(nodes[1:], 0.0), np.nextafter(nodes, np.inf)])
    x = np.concatenate([rng.uniform(0.0, L, 4000), nodes, near,
                        rng.randint(0, Ng - 1, 500) * dx])
    x = x[(x >= 0.0) & (x < L)]
    return x


def gen_pypic():
    p = refshim.load("pypic")
    rng = np.random.RandomState(11)
    out = {}
    for tag, Ng, dx in (("a", 64, 25.850471), ("b", 200, 1.0e-5), ("c", 37, 0.3)):
        L = Ng * dx
        x = adversarial_positions(Ng + 1, dx, L, rng)
        x = x[x < L * (1 - 1e-12)]
        N = len(x)
        F = rng.normal(size=Ng)
        q = -np.ones(N) * p.e
        v = rng.normal(0, 1e5, N)
        p2c = 5170.09
        out[f"{tag}_Ng"] = Ng; out[f"{tag}_dx"] = dx; out[f"{tag}_x"] = x
        out[f"{tag}_F"] = F; out[f"{tag}_v"] = v; out[f"{tag}_q"] = q; out[f"{tag}_p2c"] = p2c
        out[f"{tag}_interp"] = p.interpolate_p(F, x, Ng, N, dx)
        out[f"{tag}_j"] = p.weight_current_p(x, q, v, p2c, Ng, N, dx)
        out[f"{tag}_rho"] = p.weight_density_p(x, q, p2c, Ng, N, dx)
        out[f"{tag}_smooth"] = p.smooth_field_p(F)
        out[f"{tag}_diff"] = p.differentiate_p(F, dx, Ng)
        idx = 1. / dx
        out[f"{tag}_iL"] = (x * idx).astype(np.int64)
        out[f"{tag}_iR"] = ((x * idx + 1) % Ng).astype(np.int64)
        rho = out[f"{tag}_rho"]
        phi = p.solve_poisson_p(dx, Ng, rho, np.zeros(Ng))
        out[f"{tag}_phi"] = phi - np.max(phi)
    np.savez_compressed(os.path.join(GOLD, "pypic_kernels.npz"), **out)

    # whole Picard push: initialize_p('landau-damping') + 3 steps of particle_push_p
    np.random.seed(1)
    N, Ng = 20000, 64
    density, Kp, pert = 1e5, 1, 0.8
    Te, Ti = 100.0 * 11600., 0.1 * 11600.
    L = 22.0 * np.sqrt(p.kb * Te * p.epsilon0 / p.e / p.e / density)
    dx = L / float(Ng)
    X = np.linspace(0.0, L, Ng + 1)
    dt, tol, maxiter = 1e-5, 1e-3, 20
    m, q, x0, v0, kBTe, kBTi, growth, K, p2c, wp, invwp, LD = p.initialize_p(
        'landau-damping', N, density, Kp, pert, dx, Ng, Te, Ti, L, X)
    rho0 = p.weight_density_p(x0, q, p2c, Ng, N, dx)
    j0 = p.weight_current_p(x0, q, v0, p2c, Ng, N, dx)
    phi0 = p.solve_poisson_p(dx, Ng, rho0, np.zeros(Ng))
    phi0 = phi0 - np.max(phi0)
    E0 = -p.differentiate_p(phi0, dx, Ng)
    o = dict(N=N, Ng=Ng, L=L, dx=dx, dt=dt, tol=tol, maxiter=maxiter, p2c=p2c,
             x0=x0.copy(), v0=v0.copy(), E0=E0.copy(), j0=j0.copy(), rho0=rho0, phi0=phi0)
    ks = []
    for t in range(3):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            x1, v1, E1, j1 = p.particle_push_p(x0, v0, q, m, E0, j0, N, Ng, p2c, dx, dt, L, tol, maxiter)
        ks.append(int(re.search(r"Iterations:\s+(\d+)", buf.getvalue()).group(1)))
        E0, x0, v0, j0 = E1, x1, v1, j1
        o[f"x_{t}"] = x1.copy(); o[f"v_{t}"] = v1.copy(); o[f"E_{t}"] = E1.copy(); o[f"j_{t}"] = j1.copy()
    o["iters"] = np.array(ks)
    np.savez_compressed(os.path.join(GOLD, "pypic_push.npz"), **o)
    print("pypic golden done, iters", ks)


def gen_pypic_full():
    """BASELINE config 1(a): pypic.main's own literals (pypic.py:846-860: landau-damping, N = 1 000 000,
    Ng = 200) -- initialize_p from seed 1, then three steps of the reference's particle_push_p.  The
    particle arrays are 8 MB each, so the file keeps every STRIDE-th particle (plus the fields, which
    are complete); tests rebuild the full initial state with the drop-in's initialize_p from the same
    seed and are pinned to it by the subsample."""
    p = refshim.load("pypic")
    np.random.seed(1)
    STRIDE = 997
    N, Ng = 1000000, 200
    density, Kp, pert = 1e5, 1, 0.8
    Te, Ti = 100.0 * 11600., 0.1 * 11600.
    L = 22.0 * np.sqrt(p.kb * Te * p.epsilon0 / p.e / p.e / density)
    dx = L / float(Ng)
    X = np.linspace(0.0, L, Ng + 1)
    dt, tol, maxiter = 1e-5, 1e-3, 20
    m, q, x0, v0, kBTe, kBTi, growth, K, p2c, wp, invwp, LD = p.initialize_p(
        'landau-damping', N, density, Kp, pert, dx, Ng, Te, Ti, L, X)
    rho0 = p.weight_density_p(x0, q, p2c, Ng, N, dx)
    j0 = p.weight_current_p(x0, q, v0, p2c, Ng, N, dx)
    phi0 = p.solve_poisson_p(dx, Ng, rho0, np.zeros(Ng))
    phi0 = phi0 - np.max(phi0)
    E0 = -p.differentiate_p(phi0, dx, Ng)
    o = dict(N=N, Ng=Ng, L=L, dx=dx, dt=dt, tol=tol, maxiter=maxiter, p2c=p2c, stride=STRIDE, seed=1,
             x0_sub=x0[::STRIDE].copy(), v0_sub=v0[::STRIDE].copy(), E0=E0.copy(), j0=j0.copy(), rho0=rho0, phi0=phi0,
             x0_sum=np.sum(x0), v0_sumsq=np.sum(v0 * v0))
    ks = []
    for t in range(3):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            x1, v1, E1, j1 = p.particle_push_p(x0, v0, q, m, E0, j0, N, Ng, p2c, dx, dt, L, tol, maxiter)
        ks.append(int(re.search(r"Iterations:\s+(\d+)", buf.getvalue()).group(1)))
        E0, x0, v0, j0 = E1, x1, v1, j1
        o[f"x_{t}"] = x1[::STRIDE].copy(); o[f"v_{t}"] = v1[::STRIDE].copy(); o[f"E_{t}"] = E1.copy(); o[f"j_{t}"] = j1.copy()
        o[f"xsum_{t}"] = np.sum(x1); o[f"vsumsq_{t}"] = np.sum(v1 * v1)
    o["iters"] = np.array(ks)
    np.savez_compressed(os.path.join(GOLD, "pypic_push_1e6.npz"), **o)
    print("pypic 1e6 golden done, iters", ks)


def gen_dd_kernels():
    d = refshim.load("PIC_L_DD")
    rng = np.random.RandomState(5)
    out = {}
    for tag, Ng, dx in (("a", 51, 0.00001), ("b", 130, 0.0371)):
        L = dx * (Ng - 1)
        x = adversarial_positions(Ng, dx, L, rng)
        x = x[(x > 0) & (x < L)]
        N = len(x)
        F = rng.normal(size=Ng)
        h = N // 2
        q = np.concatenate([-np.ones(h), np.ones(N - h)]) * d.e
        v = rng.normal(0, 1e6, N)
        active = np.ones(N)
        sel = rng.uniform(size=N)
        active[sel < 0.05] = 0
        active[(sel >= 0.05) & (sel < 0.1)] = -1
        p2c, dt = 1.25e11, 1e-12
        out[f"{tag}_Ng"] = Ng; out[f"{tag}_dx"] = dx; out[f"{tag}_x"] = x; out[f"{tag}_F"] = F
        out[f"{tag}_q"] = q; out[f"{tag}_v"] = v; out[f"{tag}_active"] = active
        out[f"{tag}_p2c"] = p2c; out[f"{tag}_dt"] = dt
        out[f"{tag}_interp"] = np.array([d.interpolateField(F, xi, Ng, dx) for xi in x])
        out[f"{tag}_idx"] = np.floor(x / dx).astype(np.int64)
        out[f"{tag}_j"] = d.weightCurrents(x, q, v, p2c, Ng, N, dx, dt, active)
        out[f"{tag}_rho"] = d.weightDensities(x, q, p2c, Ng, N, dx, active)
        out[f"{tag}_diff"] = d.differentiateField(F, dx, Ng)
        out[f"{tag}_int"] = d.integrateField(F, dx, Ng)
        out[f"{tag}_smooth"] = d.smoothField(F)
    np.savez_compressed(os.path.join(GOLD, "dd_kernels.npz"), **out)
    print("dd kernels golden done")


class _Recorder:
    """Reads back the arrays the reference hands to plt.plot / plt.scatter."""

    def __init__(self, plt):
        self.plt = plt

    def calls(self, name):
        return [c for c in getattr(self.plt, name).call_args_list]


def _run_main_with_literals(modname, func, args, literals):
    """exec the (shimmed) reference module source with hard-coded literals of the
    driver replaced, run func(*args) in a scratch cwd, capture stdout + files."""
    src = refshim._py2_fix(refshim._read(modname + ".py"))
    for pat, rep in literals:
        src, n = re.subn(pat, rep, src)
        assert n >= 1, pat
    refshim._install_stubs()
    import matplotlib.pyplot as plt
    plt.reset_mock()
    mod = refshim._exec_module(modname + "_lit", src, modname + ".py")
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "plots"))
    buf = io.StringIO()
    os.chdir(tmp)
    try:
        with contextlib.redirect_stdout(buf):
            getattr(mod, func)(*args)
    finally:
        os.chdir(cwd)
    return mod, tmp, buf.getvalue(), plt


def gen_dd_main(tag, N, Ng, T, seed=1, gamma=None, vion_after=None):
    """gamma: literal override of the thermostat probability (PIC_L_DD.py:331, shipped as 0.0);
    vion_after: literal override of the `t > 2000` threshold of the vionout tally (:497,:502)."""
    np.random.seed(seed)
    lits = [(r"\n\tN = 40000\n", f"\n\tN = {N}\n"), (r"\n\tNg = 51\n", f"\n\tNg = {Ng}\n")]
    if gamma is not None:
        lits.append((r"\n\tgamma = 0\.0\n", f"\n\tgamma = {gamma!r}\n"))
    if vion_after is not None:
        lits.append((r"t > 2000", f"t > {int(vion_after)}"))
    mod, tmp, out, plt = _run_main_with_literals("PIC_L_DD", "main_i", (T, 1), lits)
    iters = np.array([int(s) for s in re.findall(r"Iterations:\s+(\d+)", out)])
    resid = np.array([float(s) for s in re.findall(r"\nr:\s+(\S+)", out)])
    E0 = np.loadtxt(os.path.join(tmp, "E0.txt"))
    jb = np.loadtxt(os.path.join(tmp, "jb.txt"))
    # recorded plot arrays: per plotted step plt.plot(X,j0), plt.plot(X,phih), plt.plot(X,E0)
    plots = [c.args[1] for c in plt.plot.call_args_list]
    j_series = np.array(plots[0::3]); phi_series = np.array(plots[1::3]); E_series = np.array(plots[2::3])
    sc = plt.scatter.call_args_list
    xi_series = np.array([c.args[0] for c in sc[0::2]])   # ions x0[N/2:]
    xe_series = np.array([c.args[0] for c in sc[1::2]])   # electrons x0[:N/2]
    ei_series = np.array([c.args[1] for c in sc[0::2]])   # sign(u)*u^2*m/2/e
    ee_series = np.array([c.args[1] for c in sc[1::2]])
    o = dict(N=N, Ng=Ng, T=T, seed=seed, iters=iters, resid=resid, E0_final=E0, jbias=jb,
             j_series=j_series, phi_series=phi_series, E_series=E_series)
    if gamma is not None:
        o["gamma"] = gamma
    if vion_after is not None:
        o["vion_after"] = vion_after
        o["vionout"] = np.atleast_1d(np.loadtxt(os.path.join(tmp, "vionout.txt")))
    o["next_uniform"] = np.random.uniform()          # pins the number of legacy-stream words the run consumed
    if N <= 4000:
        o.update(xi_series=xi_series, xe_series=xe_series, ei_series=ei_series, ee_series=ee_series)
    else:
        o.update(xi_last=xi_series[-1][::20], xe_last=xe_series[-1][::20],
                 ei_last=ei_series[-1][::20], ee_last=ee_series[-1][::20])
    np.savez_compressed(os.path.join(GOLD, f"dd_main_{tag}.npz"), **o)
    print("dd main golden", tag, "iters", iters)


def gen_pic_l():
    l = refshim.load("PIC_L")
    rng = np.random.RandomState(7)
    out = {}
    Ng, dx = 200, 0.02
    L = dx * (Ng - 1)
    x = adversarial_positions(Ng + 1, dx, L + dx, rng)
    x = x[x < (L + dx) * (1 - 1e-12)]
    N = len(x)
    q = -np.ones(N) * l.e
    m = np.ones(N) * l.me
    v = rng.normal(0, 1e6, N)
    p2c = 4.0e5
    E = rng.normal(size=Ng + 1)
    out.update(Ng=Ng, dx=dx, x=x, v=v, q=q, m=m, p2c=p2c, E=E)
    out["interp"] = np.array([l.interpolateFieldPeriodic(E, xi, Ng, dx) for xi in x])
    out["rho"] = l.weightDensitiesPeriodic(x, q, p2c, Ng, N, dx)
    out["j"] = l.weightCurrentsPeriodic(x, q, v, p2c, Ng, N, dx)
    phi = l.solvePoissonPeriodicElectronsNeutralized(dx, Ng, out["rho"], 1.0, 1e-3, 20, np.zeros(Ng + 1))
    out["phi"] = phi - np.max(phi)
    out["dphi"] = l.differentiateFieldPeriodic(out["phi"], dx, Ng)
    xo, vo = l.pushParticlesExplicit(x, v, q, m, N, Ng, 1e-9, dx, E)
    out["xout"] = xo; out["vout"] = vo
    xb, vb = l.applyBoundaryConditionsPeriodic(xo, vo, m, N, L, dx, 1.0)
    out["xbc"] = xb
    np.savez_compressed(os.path.join(GOLD, "l_kernels.npz"), **out)

    # whole explicit loop, literal override N=100000 -> 6000
    np.random.seed(1)
    mod, tmp, txt, plt = _run_main_with_literals(
        "PIC_L", "main", (12, 1), [(r"\n\tN = 100000\n", "\n\tN = 6000\n")])
    EE = np.loadtxt(os.path.join(tmp, "plots", "E2.txt"))
    plots = [c.args[1] for c in plt.plot.call_args_list]
    E_series = np.array(plots[1::2])    # plot(X,j) then plot(X,E) per step
    sc = plt.scatter.call_args_list
    np.savez_compressed(os.path.join(GOLD, "l_main.npz"), N=6000, T=12, EE=EE, E_series=E_series,
                        x_init=np.array(sc[0].args[0]), vn_init=np.array(sc[0].args[1]))
    print("PIC_L golden done")


def gen_gc():
    g = refshim.load("pygcpic")
    e, mp = g.e, g.mp
    out = {}
    # --- survey golden vectors (Boris / GC), executed again here
    B = np.array([2 * np.cos(86 * np.pi / 180), 2 * np.sin(86 * np.pi / 180), 0.0])
    pt = g.Particle(mp, 1, 1.0, 1.0, 1, B0=B.copy(), E0=np.array([1000.0, 0.0, 0.0]))
    pt.r[:] = [1e-4, 0, 0, 1e4, 2e4, -3e4, 0]
    pt.push_6D(1e-10)
    out["boris_one"] = pt.r.copy()
    # --- batch of random particles through gather / boris / transforms / GC RK4
    rng = np.random.RandomState(3)
    ng, Lg = 150, 0.0123
    grid = g.Grid(ng, Lg, 60. * 11600.)
    grid.E[:] = rng.normal(0, 5e4, ng)
    N = 400
    r0 = np.zeros((N, 7))
    r0[:, 0] = rng.uniform(0, Lg, N)
    r0[:20, 0] = np.arange(20) * grid.dx  # node aligned
    r0[:, 1:3] = rng.normal(0, 1e-4, (N, 2))
    r0[:, 3:6] = rng.normal(0, 7e4, (N, 3))
    cs = rng.choice([1, 1, 1, 2, 0], N)
    ms = rng.choice([mp, 10.81 * mp], N)
    Eshared = np.array([0.0, 30.0, -20.0])
    gath = np.zeros(N); r_b = np.zeros((N, 7)); r_gc = np.zeros((N, 7)); r_gc2 = np.zeros((N, 7))
    r_back = np.zeros((N, 7)); a_draws = np.zeros((N, 3))
    np.random.seed(17)
    st_all = []
    for i in range(N):
        pt = g.Particle(ms[i], int(cs[i]), 1.0, 1.0, 1, B0=B.copy(), E0=Eshared.copy())
        pt.r[:] = r0[i]
        pt.interpolate_electric_field_dirichlet(grid)
        gath[i] = pt.E[0]
        pt.push_6D(1e-10)
        r_b[i] = pt.r
        if cs[i] != 0:
            pt.transform_6D_to_GC()
            r_gc[i] = pt.r
            pt.push_GC(1e-10)
            r_gc2[i] = pt.r
            st = np.random.get_state()
            a_draws[i] = np.random.uniform(0.0, 1.0, 3)
            np.random.set_state(st)
            pt.transform_GC_to_6D()
            r_back[i] = pt.r
    out.update(B=B, grid_E=grid.E.copy(), ng=ng, Lg=Lg, dx=grid.dx, r0=r0, cs=cs, ms=ms,
               Eshared=Eshared, gather=gath, r_boris=r_b, r_gc=r_gc, r_gc2=r_gc2,
               r_back=r_back, a_draws=a_draws)

    # --- deposit + n0 update over 3 calls, smooth, linear & Newton solves
    ng2, L2 = 101, 1.0e-3
    Te = 50. * 11600.
    grid = g.Grid(ng2, L2, Te)
    Np = 3000
    np.random.seed(23)
    parts = [g.Particle(mp, 1, 1e19 * L2 / Np, 10 * 11600., 1, B0=B.copy(), E0=np.zeros(3), grid=grid)
             for _ in range(Np)]
    for i in range(0, Np, 7):
        parts[i].active = 0
    for i in range(0, Np, 11):
        parts[i].charge_state = 2
    out["dep_x"] = np.array([p_.r[0] for p_ in parts])
    out["dep_cs"] = np.array([p_.charge_state for p_ in parts])
    out["dep_p2c"] = np.array([p_.p2c for p_ in parts])
    out["dep_active"] = np.array([p_.active for p_ in parts])
    out.update(dep_ng=ng2, dep_L=L2, dep_Te=Te, dep_dt=1e-10)
    n0s = []; rhos = []; ns = []; phis = []; Es = []; pold = []
    for it in range(3):
        grid.weight_particles_to_grid_boltzmann(parts, 1e-10)
        rhos.append(grid.rho.copy()); ns.append(grid.n.copy()); n0s.append(grid.n0); pold.append(grid.p_old)
        grid.smooth_rho()
        grid.reset_added_particles()
        grid.add_particles(parts[0].p2c * (it + 1))
        grid.solve_for_phi_dirichlet_boltzmann()
        phis.append(grid.phi.copy())
        grid.differentiate_phi_to_E_dirichlet()
        Es.append(grid.E.copy())
    out.update(dep_rho=np.array(rhos), dep_n=np.array(ns), dep_n0=np.array(n0s), dep_pold=np.array(pold),
               dep_phi=np.array(phis), dep_E=np.array(Es), dep_rho_smooth=grid.rho.copy())
    # linear dirichlet KAT-like on random rho
    grid = g.Grid(64, 2.0, Te)
    grid.rho[:] = np.random.RandomState(4).normal(size=64)
    out["lin_rho"] = grid.rho.copy(); out["lin_dx"] = grid.dx
    grid.solve_for_phi_dirichlet()
    out["lin_phi"] = grid.phi.copy()
    # dirichlet-neumann Newton
    with contextlib.redirect_stdout(io.StringIO()):
        gdn = g.Grid(ng2, L2, Te, bc='dirichlet-neumann')
    gdn.n[:] = ns[0]
    gdn.n0 = n0s[0]
    gdn.solve_for_phi_dirichlet_neumann_boltzmann()
    out["dn_phi"] = gdn.phi.copy(); out["dn_n"] = ns[0]; out["dn_n0"] = n0s[0]; out["dn_dx"] = gdn.dx

    # --- mini driver: the particle loop of pic_bca_aps (pygcpic.py:1486-1563)
    # without BCA / ionisation, 25 steps, reference objects, global RNG seeded.
    np.random.seed(29)
    density = 1e19
    Ti = 10. * 11600; Te = 50. * 11600
    LD = np.sqrt(g.kb * Te * g.epsilon0 / e / e / density)
    Ld = 40 * LD; ngd = 121; Nd = 2000; dt = 8e-11
    p2c = density * Ld / Nd
    source_N = Nd - 40
    grid = g.Grid(ngd, Ld, Te)
    parts = [g.Particle(mp, 1, p2c, Ti, Z=1, B0=B.copy(), E0=np.zeros(3), grid=grid) for _ in range(Nd)]
    out["drv_r_init"] = np.array([p_.r.copy() for p_ in parts])
    out.update(drv_L=Ld, drv_ng=ngd, drv_N=Nd, drv_dt=dt, drv_p2c=p2c, drv_Ti=Ti, drv_Te=Te,
               drv_source_N=source_N, drv_seed=29)
    src = g.source_distribution_6D(grid, Ti, mp)
    time = 0.
    deletion_flags = []
    n_hist = []; hits = []; ndel = []; nreact = []; n0h = []; phimax = []
    ekin = []; angs = []
    for step in range(25):
        time += dt
        for p_ in parts:
            p_.apply_BCs_dirichlet(grid)
        grid.weight_particles_to_grid_boltzmann(parts, dt)
        grid.smooth_rho()
        grid.reset_added_particles()
        grid.solve_for_phi_dirichlet_boltzmann()
        grid.differentiate_phi_to_E_dirichlet()
        nh = 0; nr = 0; ek = []; an = []
        for pi, p_ in enumerate(parts):
            if p_.is_active():
                p_.interpolate_electric_field_dirichlet(grid)
                p_.push_6D(dt)
                p_.apply_BCs_dirichlet(grid)
                if not p_.is_active() and p_.at_wall:
                    nh += 1
                    ek.append(p_.kinetic_energy / e); an.append(p_.get_angle_wrt_wall())
            else:
                if sum(1 for q_ in parts if (q_.Z == 1 and q_.is_active() and q_.charge_state > 0)) < source_N:
                    p_.reactivate(src, grid, time, p2c, mp, 1, 1)
                    p_.from_wall = 0; p_.at_wall = 0
                    nr += 1
                else:
                    deletion_flags.append(pi)
        parts = [p_ for pi, p_ in enumerate(parts) if pi not in set(deletion_flags)]
        ndel.append(len(deletion_flags)); deletion_flags = []
        n_hist.append(len(parts)); hits.append(nh); nreact.append(nr); n0h.append(grid.n0)
        phimax.append(np.max(grid.phi)); ekin.append(np.array(ek)); angs.append(np.array(an))
    out["drv_r_final"] = np.array([p_.r.copy() for p_ in parts])
    out["drv_active_final"] = np.array([p_.active for p_ in parts])
    out.update(drv_len=np.array(n_hist), drv_hits=np.array(hits), drv_ndel=np.array(ndel),
               drv_nreact=np.array(nreact), drv_n0=np.array(n0h), drv_phimax=np.array(phimax),
               drv_phi_final=grid.phi.copy(), drv_rho_final=grid.rho.copy(),
               drv_ekin=np.concatenate(ekin), drv_ang=np.concatenate(angs))
    np.savez_compressed(os.path.join(GOLD, "gc.npz"), **out)
    print("pygcpic golden done; lens", n_hist[-5:], "hits", sum(hits), "del", sum(ndel), "react", sum(nreact))


def gen_gc_ion():
    """Monte-Carlo ionisation (pygcpic.py:350-458) inside the particle loop of pic_bca_aps
    (pygcpic.py:1496-1549: push, walls, ionisation attempts, mid-domain exit of wall-born
    particles, reactivate-or-delete) driven with the reference's own objects.  Neutral hydrogen
    and boron in charge states 0..2 are mixed into the ion population; their p2c is scaled so
    that the ionisation probabilities are O(0.1)."""
    g = refshim.load("pygcpic")
    e, mp = g.e, g.mp
    out = {}
    B = np.array([2 * np.cos(86 * np.pi / 180), 2 * np.sin(86 * np.pi / 180), 0.0])
    np.random.seed(41)
    density = 1e19
    Ti = 10. * 11600; Te = 50. * 11600
    LD = np.sqrt(g.kb * Te * g.epsilon0 / e / e / density)
    Ld = 40 * LD; ngd = 121; dt = 8e-11
    n_ion, n_h0, n_b = 1500, 300, 300
    Nd = n_ion + n_h0 + n_b
    p2c = density * Ld / n_ion
    p2c_n = p2c * 2e-4                      # neutrals / boron: probabilities of order 0.1
    source_N = n_ion - 30
    grid = g.Grid(ngd, Ld, Te)
    kinds = np.array([0] * n_ion + [1] * n_h0 + [2] * n_b)
    np.random.shuffle(kinds)
    parts = []
    for kd in kinds:
        if kd == 0:
            p_ = g.Particle(mp, 1, p2c, Ti, Z=1, B0=B.copy(), E0=np.zeros(3), grid=grid)
        elif kd == 1:
            p_ = g.Particle(mp, 0, p2c_n, Ti, Z=1, B0=B.copy(), E0=np.zeros(3), grid=grid)
            p_.from_wall = int(np.random.uniform() < 0.5)
        else:
            p_ = g.Particle(10.81 * mp, int(np.random.randint(0, 3)), p2c_n, Ti, Z=5, B0=B.copy(), E0=np.zeros(3), grid=grid)
            p_.from_wall = int(np.random.uniform() < 0.5)
        parts.append(p_)
    out["r_init"] = np.array([p_.r.copy() for p_ in parts])
    out["cs_init"] = np.array([float(p_.charge_state) for p_ in parts]); out["m_init"] = np.array([p_.m for p_ in parts])
    out["p2c_init"] = np.array([p_.p2c for p_ in parts]); out["Z_init"] = np.array([p_.Z for p_ in parts])
    out["from_wall_init"] = np.array([p_.from_wall for p_ in parts])
    out.update(L=Ld, ng=ngd, N=Nd, dt=dt, p2c=p2c, Ti=Ti, Te=Te, source_N=source_N, B=B)
    out["rng_state_keys"] = np.random.get_state()[1].copy(); out["rng_state_pos"] = np.random.get_state()[2]
    src = g.source_distribution_6D(grid, Ti, mp)
    time = 0.
    deletion_flags = []
    hist = dict(length=[], hits=[], ndel=[], nreact=[], nion_h=[], nion_b=[], nexit=[], n0=[], added=[])
    sink = io.StringIO()
    for step in range(20):
        time += dt
        for p_ in parts:
            p_.apply_BCs_dirichlet(grid)
        grid.weight_particles_to_grid_boltzmann(parts, dt)
        grid.smooth_rho()
        grid.reset_added_particles()
        grid.solve_for_phi_dirichlet_boltzmann()
        grid.differentiate_phi_to_E_dirichlet()
        nh = nr = nih = nib = nex = 0
        for pi, p_ in enumerate(parts):
            if p_.is_active():
                p_.interpolate_electric_field_dirichlet(grid)
                p_.push_6D(dt)
                p_.apply_BCs_dirichlet(grid)
                cs_before = p_.charge_state
                with contextlib.redirect_stdout(sink):
                    if p_.Z == 1 and p_.charge_state == 0 and p_.is_active():
                        p_.attempt_first_ionization(dt, Te, grid)
                    if p_.Z == 5 and p_.charge_state < 3 and p_.is_active():
                        p_.attempt_nth_ionization(dt, Te, grid)
                if p_.charge_state != cs_before:
                    if p_.Z == 1: nih += 1
                    else: nib += 1
                if not p_.is_active() and p_.at_wall:
                    nh += 1
                dxm = grid.length / 8
                if p_.from_wall and (grid.length / 2 - dxm < p_.x < grid.length / 2 + dxm):
                    p_.active = False
                    nex += 1
            else:
                if sum(1 for q_ in parts if (q_.Z == 1 and q_.is_active() and q_.charge_state > 0)) < source_N:
                    p_.reactivate(src, grid, time, p2c, mp, 1, 1)
                    p_.from_wall = 0; p_.at_wall = 0
                    nr += 1
                else:
                    deletion_flags.append(pi)
        dset = set(deletion_flags)
        parts = [p_ for pi, p_ in enumerate(parts) if pi not in dset]
        hist["ndel"].append(len(deletion_flags)); deletion_flags = []
        hist["length"].append(len(parts)); hist["hits"].append(nh); hist["nreact"].append(nr)
        hist["nion_h"].append(nih); hist["nion_b"].append(nib); hist["nexit"].append(nex)
        hist["n0"].append(grid.n0); hist["added"].append(grid.added_particles)
    for k_, v_ in hist.items():
        out["h_" + k_] = np.array(v_)
    out["r_final"] = np.array([p_.r.copy() for p_ in parts])
    out["cs_final"] = np.array([float(p_.charge_state) for p_ in parts]); out["Z_final"] = np.array([p_.Z for p_ in parts])
    out["active_final"] = np.array([int(p_.active) for p_ in parts])
    out["next_uniform"] = np.random.uniform()            # pins the number of draws consumed
    np.savez_compressed(os.path.join(GOLD, "gc_ion.npz"), **out)
    print("pygcpic ionisation golden done;", {k_: int(np.sum(v_)) for k_, v_ in hist.items() if k_ not in ("n0", "added")})


def gen_stub_functions():
    """The module functions no reference driver calls (PIC_L.py:48-60,83-98,146-206,261-282;
    PIC_L_DD.py:116-176; pygcpic.py:350-458), executed as they are on seeded inputs."""
    l = refshim.load("PIC_L"); d = refshim.load("PIC_L_DD"); g = refshim.load("pygcpic")
    rng = np.random.RandomState(21)
    out = {}
    Ng, dx = 64, 1e-5
    L = dx * (Ng - 1)
    N = 3000
    x = adversarial_positions(Ng, dx, L, rng); x = x[(x > 0) & (x < L)][:N]; N = len(x)
    q = np.where(rng.uniform(size=N) < 0.5, -l.e, l.e); m = np.where(q < 0, l.me, l.mp)
    v = rng.normal(0, 1e5, N)
    p2c = 3.0e9
    out.update(Ng=Ng, dx=dx, x=x, q=q, m=m, v=v, p2c=p2c)
    out["l_j"] = l.weightCurrents(x, q, v, p2c, Ng, N, dx)
    out["l_rho"] = l.weightDensities(x, q, p2c, Ng, N, dx)
    # Boltzmann-Newton solves: smooth positive charge density; the bounded variant does not converge as
    # written (its last Jacobian row does not belong to F[-1] = phi[-1]), so it is pinned after a FIXED
    # number of iterations (tol = 0 -> maxiter + 1 iterations)
    kBT = l.kb * 116000.
    X = np.arange(Ng) * dx
    rho = l.e * 1e17 * (1 + 0.05 * np.cos(2 * np.pi * X / (Ng * dx)) + 0.02 * np.sin(6 * np.pi * X / (Ng * dx)))
    phi0 = 0.3 * np.sin(np.pi * X / L)
    out.update(kBT=kBT, rho_b=rho, phi0_b=phi0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["l_phi_b1"] = l.solvePoisson(dx, Ng, rho.copy(), kBT, 0.0, 1, phi0.copy())
        out["l_phi_b3"] = l.solvePoisson(dx, Ng, rho.copy(), kBT, 0.0, 3, phi0.copy())
        out["dd_phi_b3"] = d.solvePoisson(dx, Ng, rho.copy(), kBT, 0.0, 3, phi0.copy())
    rho1 = l.e * 1e17 * (1 + 0.05 * np.cos(2 * np.pi * np.arange(Ng + 1) / (Ng + 1)))
    out["rho_p"] = rho1
    out["l_phi_p"] = l.solvePoissonPeriodic(dx, Ng, rho1.copy(), kBT, 1e-8, 20, np.zeros(Ng + 1))
    out["dd_phi_p"] = d.solvePoissonPeriodic(dx, Ng, rho1[:Ng].copy(), kBT, 1e-8, 20, np.zeros(Ng))
    # implicit push, function form (periodic gather on Ng+1 nodes)
    Eh = rng.normal(0, 1e4, Ng + 1)
    xh = rng.uniform(0, (L + dx) * (1 - 1e-9), N)
    out.update(Eh=Eh, xh=xh)
    xo, vo = l.pushParticlesImplicit(x, xh, v, q, m, N, Ng, 1e-10, dx, Eh)
    out["l_xi"] = xo; out["l_vi"] = vo
    # applyBoundaryConditions: redraws x > L or x <= 0 from the global legacy stream
    xb = x.copy(); vb = v.copy()
    k = rng.choice(N, 40, replace=False)
    xb[k[:20]] = L * (1 + rng.uniform(0, 0.1, 20)); xb[k[20:30]] = -rng.uniform(0, 1e-5, 10); xb[k[30:]] = 0.0
    out["bc_x_in"] = xb.copy(); out["bc_v_in"] = vb.copy()
    np.random.seed(5)
    xb2, vb2 = l.applyBoundaryConditions(xb, vb, m, N, L, dx, kBT)
    out["bc_x"] = xb2.copy(); out["bc_v"] = vb2.copy(); out["bc_next_uniform"] = np.random.uniform()
    # pygcpic object-level ionisation attempts on a handful of particles
    ng, Lg = 50, 5e-3
    grid = g.Grid(ng, Lg, 60. * 11600.)
    grid.n[:] = 2e15 * (1 + 0.3 * np.sin(np.arange(ng)))
    np.random.seed(9)
    rows = []
    for t in range(60):
        Z = 5 if t % 2 else 1
        cs = [0, 0, 1, 2][t % 4] if Z == 5 else 0
        pt = g.Particle(g.mp * (10.81 if Z == 5 else 1.0), cs, 2.0e6 * (1 + t % 3), 1.0, Z, grid=grid)
        pt.r[0] = (0.03 + 0.9 * ((t * 0.6180339887) % 1.0)) * Lg
        before = grid.added_particles
        which = "first" if (t % 3 or Z == 1) else "nth"
        import contextlib as _cl, io as _io
        with _cl.redirect_stdout(_io.StringIO()):
            if which == "first":
                pt.attempt_first_ionization(2e-7, 60. * 11600., grid)
            else:
                pt.attempt_nth_ionization(2e-7, 60. * 11600., grid)
        rows.append([Z, cs, pt.p2c, pt.r[0], 1.0 if which == "nth" else 0.0, float(pt.charge_state), grid.added_particles - before])
    out.update(ion_rows=np.array(rows), ion_ng=ng, ion_L=Lg, ion_n=grid.n.copy(), ion_next_uniform=np.random.uniform())
    np.savez_compressed(os.path.join(GOLD, "stub_functions.npz"), **out)
    print("stub-function golden done; ionised:", int(sum(r[5] != r[1] for r in rows)), "of", len(rows))


def main():
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["pypic", "pypicfull", "ddk", "ddm", "ddx", "l", "stubs", "gc", "gcion"]
    if "gcion" in which:
        gen_gc_ion()
    if "pypic" in which:
        gen_pypic()
    if "pypicfull" in which:
        gen_pypic_full()
    if "ddk" in which:
        gen_dd_kernels()
    if "ddm" in which:
        gen_dd_main("small", 2000, 51, 40)
        gen_dd_main("default", 40000, 51, 2)
    if "ddx" in which:
        gen_dd_main("gamma", 2000, 51, 12, gamma=0.02)
        gen_dd_main("vion", 2000, 51, 40, vion_after=5)
    if "l" in which:
        gen_pic_l()
    if "stubs" in which:
        gen_stub_functions()
    if "gc" in which:
        gen_gc()


if __name__ == "__main__":
    main()
