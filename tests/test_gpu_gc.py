"""GPU parity of the pygcpic.py path (Particle/Grid on the SoA store) against golden vectors
produced by executing the reference's classes, and against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import np_oracle as O


def relmax(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def test_particle_kernels_vs_reference_golden(golden):
    from pypic_b200.gcstore import GridDev, ParticleStore
    import torch
    g = golden("gc")
    B = g["B"]; r0 = g["r0"]; cs = g["cs"].astype(float); ms = g["ms"]
    grid = GridDev(int(g["ng"]), float(g["Lg"]), 60. * 11600.)
    assert grid.dx == float(g["dx"])
    grid.E.copy_(torch.as_tensor(g["grid_E"]))
    st = ParticleStore.from_arrays(r0, cs, ms, 1.0, B=B, Eyz=g["Eshared"][1:])
    # mirrored gather (node-aligned positions included): bit-exact
    assert np.array_equal(st.gather(grid), g["gather"])
    # Boris push: bit-exact (same operation order, no FMA contraction)
    hits = st.push_6D(1e-10, grid)
    assert np.array_equal(st.r_host(), g["r_boris"])
    exp_hit = (g["r_boris"][:, 0] < 0) | (g["r_boris"][:, 0] > float(g["Lg"]))
    assert hits == int(exp_hit.sum())
    assert np.array_equal(st.flags_host()["active"], np.where(exp_hit, 0, 1))
    # GC transforms / RK4 on the charged particles only (neutrals divide by wc = 0)
    ch = cs != 0
    st2 = ParticleStore.from_arrays(g["r_boris"][ch], cs[ch], ms[ch], 1.0, B=B, Eyz=g["Eshared"][1:])
    st2.transform_6D_to_GC()
    assert relmax(st2.r_host(), g["r_gc"][ch]) < 1e-15
    st3 = ParticleStore.from_arrays(g["r_gc"][ch], cs[ch], ms[ch], 1.0, B=B, Eyz=g["Eshared"][1:])
    # the reference's E[0] is the value gathered BEFORE the Boris push (shared Particle.E);
    # reproduce by giving the kernel a grid whose gather at the GC position returns it: use
    # the oracle path for E_x instead -- push with a uniform field per particle is not
    # expressible, so compare against the oracle evaluated with the kernel's own gather.
    Ex = st3.gather(grid)
    Evec = np.stack([Ex, np.full(ch.sum(), g["Eshared"][1]), np.full(ch.sum(), g["Eshared"][2])], 1)
    ref = O.gc_push_GC(g["r_gc"][ch], Evec, B, cs[ch], ms[ch], 1e-10)
    st3.push_GC(1e-10, grid)
    assert relmax(st3.r_host(), ref) < 1e-14
    st4 = ParticleStore.from_arrays(g["r_gc2"][ch], cs[ch], ms[ch], 1.0, B=B)
    st4.transform_GC_to_6D(g["a_draws"][ch])
    assert relmax(st4.r_host(), g["r_back"][ch]) < 1e-13
    st.check()


def test_gc_rk4_reference_vector(golden):
    """SURVEY.md P5/P6 golden vectors (executed reference): 6D->GC then one RK4 step."""
    from pypic_b200.gcstore import GridDev, ParticleStore
    import torch
    B = np.array([2 * np.cos(86 * np.pi / 180), 2 * np.sin(86 * np.pi / 180), 0.0])
    r = np.array([[1e-4, 0, 0, 1e4, 2e4, -3e4, 0]])
    st = ParticleStore.from_arrays(r, 1.0, O.mp, 1.0, B=B)
    st.transform_6D_to_GC()
    exp_gc = [1e-4, -1.7473996672903433e-24, 7.16472670814264e-24, 20648.845742637743, 4.0648850826489377e-19, -30000.0, 0]
    assert relmax(st.r_host()[0], exp_gc) < 1e-15
    grid = GridDev(11, 1e-3, 1.0)
    grid.E.fill_(1000.0)                      # uniform E_x = 1000 as in the survey probe
    st.push_GC(1e-10, grid)
    exp = [1.0014403906658945e-4, 2.0598546192239223e-6, 4.987820251299122e-8, 20648.845742684232,
           4.0648850826489377e-19, -30000.0, 1e-10]
    assert relmax(st.r_host()[0], exp) < 1e-13


def test_grid_kernels_vs_reference_golden(golden):
    from pypic_b200.gcstore import GridDev, ParticleStore
    from pypic_b200 import ops
    import torch
    g = golden("gc")
    ng = int(g["dep_ng"]); L = float(g["dep_L"]); Te = float(g["dep_Te"]); dt = float(g["dep_dt"])
    grid = GridDev(ng, L, Te)
    N = len(g["dep_x"])
    r = np.zeros((N, 7)); r[:, 0] = g["dep_x"]
    st = ParticleStore.from_arrays(r, g["dep_cs"].astype(float), O.mp, g["dep_p2c"], active=g["dep_active"])
    for it in range(3):
        grid.weight_particles_to_grid_boltzmann(st, dt)
        assert relmax(grid.rho.cpu().numpy(), g["dep_rho"][it]) < 1e-13
        assert relmax(grid.n.cpu().numpy(), g["dep_n"][it]) < 1e-13
        assert abs(grid.n0 - g["dep_n0"][it]) <= 1e-11 * abs(grid.n0)
        grid.smooth_rho()
        grid.reset_added_particles()
        grid.add_particles(float(g["dep_p2c"][0]) * (it + 1))
        grid.solve_for_phi_dirichlet_boltzmann()
        phi = grid.phi.cpu().numpy()
        # the reference's Newton step is bicgstab at default rtol: documented phi tolerance
        assert np.max(np.abs(phi - g["dep_phi"][it])) < 2e-5 * max(1.0, np.max(np.abs(phi)))
        # and the exact Newton fixed point agrees with the oracle's exact-step restatement
        pr, _ = O.gc_solve_for_phi_dirichlet_boltzmann(grid.rho.cpu().numpy(), grid.n0, Te, grid.dx)
        assert np.max(np.abs(phi - pr)) < 1e-9 * max(1.0, np.max(np.abs(pr)))
        grid.phi.copy_(torch.as_tensor(g["dep_phi"][it]))       # continue from the reference's phi
        grid.differentiate_phi_to_E_dirichlet()
        assert np.array_equal(grid.E.cpu().numpy(), g["dep_E"][it])
    assert relmax(grid.rho.cpu().numpy(), g["dep_rho_smooth"]) < 1e-13
    grid.check()
    # doctest KAT: Grid(5,4.0): rho=1 -> phi=[0,1.5,2,1.5,0]
    assert np.allclose(ops.poisson_dirichlet(np.ones(5), 1.0), [0, 1.5, 2, 1.5, 0], rtol=0, atol=1e-14)
    assert relmax(ops.poisson_dirichlet(g["lin_rho"], float(g["lin_dx"])), g["lin_phi"]) < 1e-11
    # neutral plasma -> phi == 0 (doctests pygcpic.py:1013-1019, 1070-1076)
    phi, _ = ops.newton_boltzmann(np.ones(5), None, 1.0, 1.0 / O.e, 1.0, 0, 1e-9, 1000)
    assert np.all(np.abs(phi) < 1e-12)
    phi, _ = ops.newton_boltzmann(np.ones(5) / O.e * O.epsilon0, np.zeros(5), 1.0, 1.0 / O.e * O.epsilon0, 1.0, 1, 1e-3, 100)
    assert np.all(np.abs(phi) < 1e-12)
    # Dirichlet-Neumann Newton on the golden density
    phi, it = ops.newton_boltzmann(g["dn_n"], np.zeros(ng), float(g["dn_dx"]), float(g["dn_n0"]), Te, 1, 1e-3, 100)
    assert np.max(np.abs(phi - g["dn_phi"])) < 1e-6 * max(1.0, np.max(np.abs(phi)))


def test_decide_rule_and_compaction():
    from pypic_b200.gcstore import ParticleStore
    import torch
    rs = np.random.RandomState(8)
    N = 100000
    active_entry = rs.uniform(size=N) > 0.03
    active_after = active_entry & (rs.uniform(size=N) > 0.02)
    src_entry = rs.uniform(size=N) > 0.1
    src_after = src_entry | (rs.uniform(size=N) > 0.95)
    ce = (active_entry & src_entry); ca = (active_after & src_after)
    source_N = int(ce.sum()) - 500
    react, dele = O.gc_particle_loop_decisions(active_entry, active_after, src_entry, src_after, source_N)
    st = ParticleStore(N)
    dev = st.dev
    t8 = lambda a: torch.as_tensor(a.astype(np.int8), device=dev)
    dec, nr, nd = st.decide(t8(~active_entry), t8(ce), t8(ca), source_N)
    d = dec.cpu().numpy()
    assert np.array_equal(d == 1, react) and np.array_equal(d == 2, dele)        # bit-exact decisions
    assert nr == react.sum() and nd == dele.sum() and nr > 0 and nd > 0
    # stable compaction: survivors keep their order
    r = np.zeros((N, 7)); r[:, 0] = np.arange(N)
    st = ParticleStore.from_arrays(r, 1.0, 1.0, 1.0)
    removed = st.compact(dec)
    assert removed == nd and st.N == N - nd
    assert np.array_equal(st.r_host()[:, 0], np.arange(N)[~dele])


def test_mini_driver_vs_reference_golden(golden):
    """The particle loop of pygcpic.pic_bca_aps (pygcpic.py:1486-1563, without BCA and
    ionisation) run by the reference's objects vs the SoA store on the GPU: same seed, same
    legacy-MT19937 draw order, identical integer outcomes per step."""
    from pypic_b200.gcstore import GridDev, ParticleStore
    import torch
    g = golden("gc")
    Ld = float(g["drv_L"]); ngd = int(g["drv_ng"]); Nd = int(g["drv_N"]); dt = float(g["drv_dt"])
    p2c = float(g["drv_p2c"]); Ti = float(g["drv_Ti"]); Te = float(g["drv_Te"]); source_N = int(g["drv_source_N"])
    B = g["B"]
    np.random.seed(int(g["drv_seed"]))
    vth = np.sqrt(O.kb * Ti / O.mp)
    r = np.zeros((Nd, 7))
    for i in range(Nd):                      # Particle._initialize_6D draw order, pygcpic.py:299-301
        r[i, 0] = np.random.uniform(0.0, Ld)
        r[i, 3:6] = np.random.normal(0.0, vth, 3) + 0.
    assert np.array_equal(r, g["drv_r_init"])
    grid = GridDev(ngd, Ld, Te)
    st = ParticleStore.from_arrays(r, 1.0, O.mp, p2c, Z=1, B=B)
    time = 0.
    lens, hits, ndel, nreact, n0h, phimax, ek, ang = [], [], [], [], [], [], [], []
    for step in range(25):
        time += dt
        st.apply_BCs_dirichlet(grid)
        grid.weight_particles_to_grid_boltzmann(st, dt)
        grid.smooth_rho()
        grid.reset_added_particles()
        grid.solve_for_phi_dirichlet_boltzmann()
        grid.differentiate_phi_to_E_dirichlet()
        inactive_entry = (st.active[:st.N] != 1).to(torch.int8)
        ce = st.source_ion_flags(1)[:st.N].contiguous()
        h = st.push_6D(dt, grid)
        ke, an, _ = st.wall_hit_tallies()
        ca = st.source_ion_flags(1)[:st.N].contiguous()
        dec, nr, nd = st.decide(inactive_entry, ce, ca, source_N)
        ridx = torch.nonzero(dec[:st.N] == 1).flatten().cpu().numpy()
        rn = np.zeros((len(ridx), 7))
        for j in range(len(ridx)):           # source_distribution_6D draw order, pygcpic.py:749-752
            x = np.random.normal(Ld / 2, Ld / 12.0)
            x %= Ld
            rn[j, 0] = x
            rn[j, 3:6] = np.random.normal(0.0, vth, 3) + 0.
        st.reactivate(ridx, rn, p2c, O.mp, 1, 1, time, grid)
        st.compact(dec)
        lens.append(st.N); hits.append(h); ndel.append(nd); nreact.append(nr)
        n0h.append(grid.n0); phimax.append(float(grid.phi.max().item())); ek.append(ke); ang.append(an)
    st.check(); grid.check()
    assert np.array_equal(lens, g["drv_len"]) and np.array_equal(hits, g["drv_hits"])        # exact integers
    assert np.array_equal(ndel, g["drv_ndel"]) and np.array_equal(nreact, g["drv_nreact"])
    assert relmax(n0h, g["drv_n0"]) < 1e-9
    assert relmax(phimax, g["drv_phimax"]) < 1e-4          # bicgstab tolerance of the reference
    assert np.array_equal(st.flags_host()["active"], g["drv_active_final"])
    assert relmax(st.r_host(), g["drv_r_final"]) < 1e-6
    assert relmax(np.concatenate(ek), g["drv_ekin"]) < 1e-5
    assert relmax(np.concatenate(ang), g["drv_ang"]) < 1e-5


@pytest.mark.parametrize("ng", [300, 6000])       # 6000 nodes: the field tile leaves room for a 2-stage ring only
def test_uniform_fused_boris_matches_v1_kernels(ng):
    """gc_push_boris_v2_k (species-uniform store, TMA ring, fused n deposit) against the v1
    push + separate weight kernels on the same inputs: r, flags and hit counts bit-identical,
    n and rho to 1e-13.  Inputs include inactive slots, wall hits on both sides, node-aligned
    and guard-band positions, and a tail that is not a whole chunk."""
    from pypic_b200.gcstore import GridDev, ParticleStore
    rs = np.random.RandomState(5)
    Lg = 3e-3; Te = 60 * 11600.; dt = 2e-9 * 300 / ng
    N = 3 * 16384 + 517
    dx = Lg / (ng - 1)
    x = np.sort(rs.uniform(0, Lg, N))
    j = rs.choice(N, 3000, replace=False)
    cells = rs.randint(1, ng - 1, 3000)
    x[j[:800]] = cells[:800] * dx
    x[j[800:1600]] = np.nextafter(cells[800:1600] * dx, 0.0)
    x[j[1600:2200]] = cells[1600:2200] * dx * (1 + 1e-9)
    x[j[2200:2600]] = rs.uniform(0, 2e-2 * dx, 400)
    x[j[2600:]] = Lg - rs.uniform(0, 2e-2 * dx, 400)
    r = np.zeros((N, 7))
    r[:, 0] = x; r[:, 1:3] = rs.normal(0, 1e-4, (N, 2)); r[:, 3:6] = rs.normal(0, 7e4, (N, 3)); r[:, 6] = rs.uniform(0, 1e-8, N)
    active = np.ones(N, dtype=np.int8); active[rs.choice(N, 700, replace=False)] = 0
    B = (2 * np.cos(1.5), 2 * np.sin(1.5), 0.)
    E = rs.normal(0, 5e4, ng)
    p2c = 3.1e9
    res = {}
    for fused in (False, True):
        grid = GridDev(ng, Lg, Te)
        grid.E.copy_(__import__("torch").as_tensor(E))
        st = ParticleStore.from_arrays(r, 1.0, O.mp, p2c, Z=1, active=active, B=B)
        st.FUSED_MIN = 0 if fused else 10 ** 12
        hits = st.push_6D(dt, grid, deposit=fused)
        if fused:
            assert grid.have_fused_n
            grid.finish_fused_deposit(1.0, dt)
        else:
            st.apply_BCs_dirichlet(grid)
            grid.weight_particles_to_grid_boltzmann(st, dt)
        st.check(); grid.check()
        res[fused] = (st.r_host(), st.flags_host(), hits, st.hit_flag[:N].cpu().numpy(), grid.n.cpu().numpy(),
                      grid.rho.cpu().numpy(), grid.n0)
    a, b = res[False], res[True]
    assert np.array_equal(a[0], b[0])
    for kf in ("active", "at_wall", "from_wall"):
        assert np.array_equal(a[1][kf], b[1][kf])
    assert a[2] == b[2] and a[2] > 200
    assert np.array_equal(a[3], b[3])
    assert relmax(b[4], a[4]) < 1e-13 and relmax(b[5], a[5]) < 1e-13
    assert abs(b[6] - a[6]) <= 1e-13 * abs(a[6])


@pytest.mark.parametrize("ng,lean", [(300, False), (4097, False), (300, True)])
def test_mixed_species_fused_boris_matches_v1_kernels(ng, lean):
    """gc_push_boris_mix_k (several species in one list: per-particle charge_state, m, p2c in the TMA
    ring, n and rho deposits fused) against the v1 push + separate weight kernels: r, flags, hit counts
    bit-identical; n, rho and the Boltzmann n0 to 1e-13.  Hydrogen ions, neutrals (charge 0), B+, B2+
    with different weights; inactive slots, wall hits, node-aligned positions, a ragged tail; several
    steps so that particles leave their deposit windows."""
    import torch
    from pypic_b200.gcstore import GridDev, ParticleStore
    rs = np.random.RandomState(15)
    Lg = 3e-3; Te = 60 * 11600.; dt = 2e-9 * 300 / ng
    N = 3 * 16384 + 517
    dx = Lg / (ng - 1)
    x = np.sort(rs.uniform(0, Lg, N))
    j = rs.choice(N, 2400, replace=False)
    cells = rs.randint(1, ng - 1, 2400)
    x[j[:800]] = cells[:800] * dx
    x[j[800:1600]] = np.nextafter(cells[800:1600] * dx, 0.0)
    x[j[1600:2000]] = rs.uniform(0, 2e-2 * dx, 400)
    x[j[2000:]] = Lg - rs.uniform(0, 2e-2 * dx, 400)
    r = np.zeros((N, 7))
    r[:, 0] = x; r[:, 1:3] = rs.normal(0, 1e-4, (N, 2)); r[:, 3:6] = rs.normal(0, 7e4, (N, 3)); r[:, 6] = rs.uniform(0, 1e-8, N)
    sp = rs.choice(4, N, p=[0.6, 0.15, 0.15, 0.1])
    cs = np.array([1., 0., 1., 2.])[sp]; m = np.array([O.mp, O.mp, 10.81 * O.mp, 10.81 * O.mp])[sp]
    p2c = np.array([3.1e9, 1.0e9, 2.0e8, 2.0e8])[sp]; Zs = np.array([1, 1, 5, 5])[sp]
    active = np.ones(N, dtype=np.int8); active[rs.choice(N, 700, replace=False)] = 0
    B = (2 * np.cos(1.5), 2 * np.sin(1.5), 0.)
    E = rs.normal(0, 5e4, ng)
    res = {}
    for fused in (False, True):
        grid = GridDev(ng, Lg, Te)
        grid.E.copy_(torch.as_tensor(E))
        st = ParticleStore.from_arrays(r, cs, m, p2c, Z=Zs, active=active, B=B)
        assert st.uniform() is None
        st.FUSED_MIN = 0 if fused else 10 ** 12
        if lean and fused:
            st.carry_yzt = False
        hits = []
        for _ in range(3):
            hits.append(st.push_6D(dt, grid, deposit=fused))
            if fused:
                assert grid.have_fused_n and grid._fused_mixed
                grid.finish_fused_deposit(1.0, dt)
            else:
                st.apply_BCs_dirichlet(grid)
                grid.weight_particles_to_grid_boltzmann(st, dt)
        st.check(); grid.check()
        res[fused] = (st.r_host(), st.flags_host(), hits, st.hit_flag[:N].cpu().numpy(), grid.n.cpu().numpy(),
                      grid.rho.cpu().numpy(), grid.n0)
    a, b = res[False], res[True]
    cols = (0, 3, 4, 5) if lean else range(7)
    for c in cols:
        assert np.array_equal(a[0][:, c], b[0][:, c]), c
    for kf in ("active", "at_wall", "from_wall"):
        assert np.array_equal(a[1][kf], b[1][kf])
    assert a[2] == b[2] and sum(a[2]) > 200
    assert np.array_equal(a[3], b[3])
    assert relmax(b[4], a[4]) < 1e-13 and relmax(b[5], a[5]) < 1e-13
    assert abs(b[6] - a[6]) <= 1e-13 * abs(a[6])


def test_species_change_promotes_the_pending_fused_deposit():
    """A species-uniform store whose fused push has deposited n (rho implied) receives re-activated
    particles of ANOTHER species: the pending deposit becomes a two-accumulator one
    (GridDev.promote_fused_to_mixed), the next push takes the mixed kernel, and n, rho equal the
    separate weight pass over the final store (pygcpic.py:871-883)."""
    import torch
    from pypic_b200.gcstore import GridDev, ParticleStore
    rs = np.random.RandomState(16)
    ng = 301; Lg = 3e-3; Te = 60 * 11600.; dt = 2e-9
    N = 2 * 16384 + 31
    r = np.zeros((N, 7)); r[:, 0] = rs.uniform(0, Lg, N); r[:, 3:6] = rs.normal(0, 7e4, (N, 3))
    active = np.ones(N, dtype=np.int8); dead = rs.choice(N, 400, replace=False); active[dead] = 0
    grid = GridDev(ng, Lg, Te); grid.E.copy_(torch.as_tensor(rs.normal(0, 5e4, ng)))
    st = ParticleStore.from_arrays(r, 1.0, O.mp, 3.1e9, Z=1, active=active, B=(0.3, 1.9, 0.))
    st.push_6D(dt, grid, deposit=True)
    assert grid.have_fused_n and not grid._fused_mixed
    idx = np.sort(dead[:150])
    r_new = np.zeros((150, 7)); r_new[:, 0] = rs.uniform(0, Lg, 150); r_new[:, 3:6] = rs.normal(0, 2e4, (150, 3))
    st.reactivate(idx, r_new, 2.0e8, 10.81 * O.mp, 2.0, 5, dt, grid)
    assert grid._fused_mixed and st.uniform() is None
    grid.finish_fused_deposit(1.0, dt)
    n1, rho1 = grid.n.cpu().numpy(), grid.rho.cpu().numpy()
    ref = GridDev(ng, Lg, Te)
    ref.weight_particles_to_grid_boltzmann(st, dt)
    assert relmax(n1, ref.n.cpu().numpy()) < 1e-13 and relmax(rho1, ref.rho.cpu().numpy()) < 1e-13
    # and the next push of the (now mixed) store deposits both accumulators itself
    st.push_6D(dt, grid, deposit=True)
    assert grid._fused_mixed
    grid.finish_fused_deposit(1.0, dt)
    st.apply_BCs_dirichlet(ref)
    ref.weight_particles_to_grid_boltzmann(st, dt)
    assert relmax(grid.n.cpu().numpy(), ref.n.cpu().numpy()) < 1e-13
    assert relmax(grid.rho.cpu().numpy(), ref.rho.cpu().numpy()) < 1e-13
    st.check(); grid.check()


def test_lean_store_boris_is_bit_identical_in_x_and_v():
    """ParticleStore.carry_yzt = False: the fused kernel streams x, vx, vy, vz only (64 B per
    particle-step).  Over several steps with wall hits: x, v, flags, hit counts and the deposited
    density bit-identical to the full store; the clock of every particle is still known (time of the
    last push for the active ones, time of death for the absorbed ones); y, z are reported as NaN."""
    import torch
    from pypic_b200.gcstore import GridDev, ParticleStore
    rs = np.random.RandomState(6)
    ng = 301; Lg = 3e-3; Te = 60 * 11600.; dt = 2e-9
    N = 4 * 16384 + 77
    r = np.zeros((N, 7))
    r[:, 0] = np.sort(rs.uniform(0, Lg, N)); r[:, 3:6] = rs.normal(0, 7e4, (N, 3))
    B = (2 * np.cos(1.5), 2 * np.sin(1.5), 0.)
    E = rs.normal(0, 5e4, ng)
    out = {}
    for lean in (False, True):
        grid = GridDev(ng, Lg, Te); grid.E.copy_(torch.as_tensor(E))
        st = ParticleStore.from_arrays(r, 1.0, O.mp, 3.1e9, Z=1, B=B)
        st.carry_yzt = not lean
        hits = []
        for s_ in range(5):
            hits.append(st.push_6D(dt, grid, deposit=True))
            grid.finish_fused_deposit(1.0, dt)
        st.check(); grid.check()
        out[lean] = (st.r_host(), st.flags_host(), hits, grid.n.cpu().numpy())
    a, b = out[False], out[True]
    assert a[2] == b[2] and sum(a[2]) > 100
    for c in (0, 3, 4, 5):
        assert np.array_equal(a[0][:, c], b[0][:, c])
    assert np.array_equal(a[1]["active"], b[1]["active"]) and np.array_equal(a[1]["at_wall"], b[1]["at_wall"])
    assert relmax(b[3], a[3]) < 1e-13                 # same contributions, merged in a different order
    assert np.all(np.isnan(b[0][:, 1:3]))
    # clocks: the full store accumulates t += dt per push; the lean one reports the same numbers
    assert np.array_equal(a[0][:, 6], b[0][:, 6])


@pytest.mark.parametrize("event_loop", [False, True])     # True: the global-event-list formulation the sharded runs use
def test_mini_driver_fused_path_vs_reference_golden(golden, event_loop):
    """pygcpic.run_sheath with the fused push+deposit path forced on (uniform store): the same
    integer outcomes per step as the reference's object loop."""
    import pygcpic as G
    g = golden("gc")
    Ld = float(g["drv_L"]); ngd = int(g["drv_ng"]); Nd = int(g["drv_N"]); dt = float(g["drv_dt"])
    p2c = float(g["drv_p2c"]); Ti = float(g["drv_Ti"]); Te = float(g["drv_Te"]); source_N = int(g["drv_source_N"])
    np.random.seed(int(g["drv_seed"]))
    host_grid = G.Grid(ngd, Ld, Te)
    parts = [G.Particle(G.mp, 1, p2c, Ti, Z=1, B0=g["B"].copy(), E0=np.zeros(3), grid=host_grid) for _ in range(Nd)]
    r = np.array([p.r for p in parts])
    grid = G.GridDev(ngd, Ld, Te)
    st = G.ParticleStore.from_arrays(r, 1.0, G.mp, p2c, Z=1, B=g["B"])
    st.FUSED_MIN = 0
    src = G.source_distribution_6D(host_grid, Ti, G.mp)
    out = G.run_sheath(grid, st, dt, 25, source_N, src, p2c, G.mp, event_loop=event_loop)
    assert np.array_equal(out["length"], g["drv_len"]) and np.array_equal(out["hits"], g["drv_hits"])
    assert np.array_equal(out["deleted"], g["drv_ndel"]) and np.array_equal(out["reactivated"], g["drv_nreact"])
    assert relmax(out["n0"], g["drv_n0"]) < 1e-9
    assert np.array_equal(st.flags_host()["active"], g["drv_active_final"])
    assert relmax(st.r_host(), g["drv_r_final"]) < 1e-6


def test_store_sort_by_cell_is_a_cell_ordered_permutation():
    import torch
    from pypic_b200.gcstore import GridDev, ParticleStore
    rs = np.random.RandomState(8)
    N, ng, Lg = 70001, 129, 2e-3
    r = rs.normal(size=(N, 7)); r[:, 0] = rs.uniform(0, Lg, N)
    cs = rs.randint(1, 3, N).astype(float); m = rs.uniform(1, 2, N) * O.mp; p2c = rs.uniform(1, 2, N) * 1e9
    Z = rs.randint(1, 5, N); act = (rs.uniform(size=N) < 0.9).astype(np.int8); aw = 1 - act
    grid = GridDev(ng, Lg, 6e5)
    st = ParticleStore.from_arrays(r, cs, m, p2c, Z=Z, active=act, at_wall=aw, B=(0, 0, 1))
    st.sort_by_cell(grid, track=True)
    perm = st.perm.cpu().numpy()
    assert np.array_equal(np.sort(perm), np.arange(N))
    assert np.array_equal(st.r_host(), r[perm])
    assert np.array_equal(st.charge_state.cpu().numpy(), cs[perm]) and np.array_equal(st.m.cpu().numpy(), m[perm])
    assert np.array_equal(st.p2c.cpu().numpy(), p2c[perm]) and np.array_equal(st.Z.cpu().numpy(), Z[perm])
    f = st.flags_host()
    assert np.array_equal(f["active"], act[perm]) and np.array_equal(f["at_wall"], aw[perm])
    cells = np.floor(st.r_host()[:, 0] / grid.dx)
    assert np.all(np.diff(cells) >= 0)
    st.sort_by_cell(grid, track=True)          # a second sort composes
    assert np.array_equal(st.r_host(), r[st.perm.cpu().numpy()])


def test_uniform_rk4_kernel_is_bit_identical_to_v1():
    """gc_push_rk4_uniform_k (drift terms hoisted, constant-divisor divisions) vs gc_push_rk4_k
    on the same guiding-centre state: every component bit-identical (incl. negative and tiny
    numerators of the constant-divisor divisions), several steps."""
    import torch
    from pypic_b200.gcstore import GridDev, ParticleStore
    rs = np.random.RandomState(13)
    N, ng, Lg = 300000, 200, 2.5e-3
    r = np.zeros((N, 7))
    r[:, 0] = rs.uniform(0.05 * Lg, 0.95 * Lg, N); r[:, 1:3] = rs.normal(0, 1e-3, (N, 2))
    r[:, 3] = rs.normal(0, 7e4, N); r[rs.choice(N, 1000), 3] *= 1e-6        # small v_par: large rho-divisions
    r[:, 4] = np.abs(rs.normal(0, 1e-17, N)); r[:, 5] = rs.normal(0, 7e4, N)
    E = rs.normal(0, 5e4, ng)
    out = {}
    for Bv, Eyz in (((2 * np.cos(1.5), 2 * np.sin(1.5), 0.), (0., 0.)), ((0.3, -1.1, 0.7), (35., -12.))):
        for uni in (True, False):
            grid = GridDev(ng, Lg, 6e5)
            grid.E.copy_(torch.as_tensor(E))
            st = ParticleStore.from_arrays(r, 1.0, O.mp, 2e9, Z=1, B=Bv, Eyz=Eyz)
            st.mode = 1
            st.RK4_UNIFORM = uni
            for _ in range(3):
                st.push_GC(1e-10, grid)
            st.check()
            out[uni] = st.r_host()
        assert np.array_equal(out[True], out[False])
        assert not np.array_equal(out[True][:, :4], r[:, :4])


def test_boris_at_baseline_config_3_size_against_the_oracle_on_a_sample():
    """BASELINE config 3 at its full size: 2e8 H+ in a field at 86 degrees to the wall normal, 4097 nodes, the fused
    lean kernel (gather + Boris + walls + density deposit) on a cell-sorted store, three steps.  A random sample of
    2e5 particles is checked BIT FOR BIT against the NumPy restatement of pygcpic.py:344-347 (mirrored gather),
    :460-507 (Boris) and :668-689 (walls) every step; over ALL particles: the wall hits the kernel counts equal the
    flags it cleared, and the deposited density integrates to the number of survivors (every particle puts
    p2c/dx * (w_l + w_r) on the grid, pygcpic.py:873-883)."""
    import torch
    from pypic_b200.gcstore import GridDev, ParticleStore
    free, _ = torch.cuda.mem_get_info()
    N = 200_000_000 if free > 60e9 else 20_000_000
    ng = 4097
    Te, Ti = 60. * 11600., 50. * 11600.
    lamD = np.sqrt(8.854e-12 * O.kb * Te / (1e19 * O.e ** 2))
    Lg = 100. * lamD * (ng - 1) / 149.
    alpha = 86. * np.pi / 180.
    B = (2. * np.cos(alpha), 2. * np.sin(alpha), 0.)
    dt = 1e-10
    dev = torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device=dev); gen.manual_seed(77)
    grid = GridDev(ng, Lg, Te)
    rs = np.random.RandomState(8)
    E = rs.normal(0, 5e4, ng)
    grid.E.copy_(torch.as_tensor(E))
    dx = grid.dx
    store = ParticleStore(N, B=B)
    store.r[0].uniform_(0., 1., generator=gen).mul_(Lg).clamp_(Lg * 1e-9, Lg * (1 - 1e-9))
    # a thin layer next to each wall so that every step absorbs particles there
    edge = N // 2000
    store.r[0][:edge].mul_(2e-3); store.r[0][edge:2 * edge].mul_(2e-3).neg_().add_(Lg)
    vth = float(np.sqrt(O.kb * Ti / O.mp))
    for c in (3, 4, 5):
        store.r[c].normal_(0., vth, generator=gen)
    p2c = Lg * 1e19 / N
    store.charge_state.fill_(1.); store.m.fill_(O.mp); store.p2c.fill_(p2c); store.Z.fill_(1)
    store.carry_yzt = False
    store.sort_by_cell(grid)
    idx = torch.randint(0, N, (200000,), generator=gen, device=dev)
    total_hits = 0
    for step in range(3):
        xb = store.r[0][idx].cpu().numpy()
        vb = np.stack([store.r[c][idx].cpu().numpy() for c in (3, 4, 5)], 1)
        ab = store.active[idx].cpu().numpy()
        hits = store.push_6D(dt, grid, deposit=True)
        grid.finish_fused_deposit(1.0, dt)
        store.check(); grid.check()
        total_hits += hits
        xa = store.r[0][idx].cpu().numpy()
        va = np.stack([store.r[c][idx].cpu().numpy() for c in (3, 4, 5)], 1)
        aa = store.active[idx].cpu().numpy()
        r = np.zeros((len(xb), 7)); r[:, 0] = xb; r[:, 3:6] = vb
        live = ab == 1
        Ex = np.zeros(len(xb)); Ex[live] = O.gc_gather_mirrored(E, xb[live], dx)
        r2 = O.gc_push_6D(r, Ex, B, 1.0, O.mp, dt)
        act2, _ = O.gc_apply_BCs_dirichlet(r2[:, 0], ab.astype(np.int64), np.zeros(len(xb), np.int64), Lg)
        assert np.array_equal(aa[live], act2[live])
        assert np.array_equal(xa[live], r2[live, 0]) and np.array_equal(va[live], r2[live, 3:6])
        assert np.array_equal(xa[~live], xb[~live]) and np.array_equal(va[~live], vb[~live])       # dead slots are not touched
        dead_all = int((store.active[:N] != 1).sum().item())
        assert dead_all == total_hits and hits > 0
        n = grid.n.cpu().numpy()
        assert abs(n.sum() * dx / p2c - (N - dead_all)) <= 1e-9 * N
    assert total_hits > 100
