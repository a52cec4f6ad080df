"""GPU parity tests of the DROP-IN MODULES (pypic, PIC_L, PIC_L_DD, pygcpic at the repository
root): the reference's own module-level API, driven the way the reference's launch scripts and
doctests drive it, compared with golden vectors produced by executing the reference
(tests/golden/*.npz, generator oracle/make_golden.py) and with the known-answer doctests the
reference ships."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@contextlib.contextmanager
def scratch_cwd(tmp_path):
    cwd = os.getcwd()
    os.makedirs(os.path.join(tmp_path, "plots"), exist_ok=True)
    os.chdir(tmp_path)
    try:
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            yield buf
    finally:
        os.chdir(cwd)


# ------------------------------------------------------------------------------- PIC_L_DD
@pytest.mark.parametrize("sort_every", [0, 2])
def test_pic_l_dd_main_i_thermostat_golden(golden, tmp_path, sort_every):
    """gamma != 0 (the reference ships gamma = 0.0; golden made with the literal overridden): the
    thermostat's uniforms and the normals of its hits interleave in the legacy stream exactly as in
    PIC_L_DD.py:419-427, the hits land on the same particles (also in a cell-sorted store) and the
    stream ends at the same word."""
    import re
    import PIC_L_DD
    g = golden("dd_main_gamma")
    N = int(g["N"]); Ng = int(g["Ng"]); T = int(g["T"])
    res = {}
    np.random.seed(int(g["seed"]))
    with scratch_cwd(str(tmp_path)) as buf:
        PIC_L_DD.main_i(T, 10, N=N, Ng=Ng, gamma=float(g["gamma"]), result=res, sort_every=sort_every)
        E0 = np.loadtxt("E0.txt"); jb = np.loadtxt("jb.txt")
    assert np.random.uniform() == float(g["next_uniform"])
    iters = np.array([int(s) for s in re.findall(r"Iterations:\s+(\d+)", buf.getvalue())])
    assert np.array_equal(iters, g["iters"])
    assert relmax(E0, g["E0_final"]) < 1e-11 and relmax(jb, g["jbias"]) < 1e-8
    h = N // 2
    assert relmax(res["x0"][:h], g["xe_series"][-1]) < 1e-11 and relmax(res["x0"][h:], g["xi_series"][-1]) < 1e-11
    ee = np.sign(res["u0"]) * res["u0"] ** 2 * 0.5 * np.where(np.arange(N) < h, O.me, O.mp) / O.e
    assert relmax(ee[:h], g["ee_series"][-1]) < 1e-10 and relmax(ee[h:], g["ei_series"][-1]) < 1e-10


@pytest.mark.parametrize("sort_every", [0, 3])
def test_pic_l_dd_main_i_vionout_golden(golden, tmp_path, sort_every):
    """vionout.txt (PIC_L_DD.py:497-503; golden made with the `t > 2000` literal lowered to 5): the
    velocities of the absorbed electrons, signed by wall, in (step, Picard iteration, index) order."""
    import PIC_L_DD
    g = golden("dd_main_vion")
    N = int(g["N"]); Ng = int(g["Ng"]); T = int(g["T"])
    np.random.seed(int(g["seed"]))
    with scratch_cwd(str(tmp_path)):
        PIC_L_DD.main_i(T, 10, N=N, Ng=Ng, vion_after=int(g["vion_after"]), sort_every=sort_every)
        vion = np.atleast_1d(np.loadtxt("vionout.txt")); E0 = np.loadtxt("E0.txt")
    assert np.random.uniform() == float(g["next_uniform"])
    assert vion.shape == g["vionout"].shape and len(vion) > 10
    assert np.array_equal(np.sign(vion), np.sign(g["vionout"]))
    assert relmax(vion, g["vionout"]) < 1e-11
    assert relmax(E0, g["E0_final"]) < 1e-11


@pytest.mark.parametrize("tag", ["small", "default"])
def test_pic_l_dd_main_i_module(golden, tag, tmp_path):
    """`import PIC_L_DD; PIC_L_DD.main_i(T, nplot)` (what run_pypic_dd.py does) with the seed the
    golden run used: iteration counts, E0.txt, jb.txt and the final particles match the reference."""
    import re
    import PIC_L_DD
    g = golden("dd_main_" + tag)
    N = int(g["N"]); Ng = int(g["Ng"]); T = int(g["T"])
    res = {}
    np.random.seed(int(g["seed"]))
    with scratch_cwd(str(tmp_path)) as buf:
        PIC_L_DD.main_i(T, 10, N=N, Ng=Ng, result=res)
        E0 = np.loadtxt("E0.txt"); jb = np.loadtxt("jb.txt")
    iters = np.array([int(s) for s in re.findall(r"Iterations:\s+(\d+)", buf.getvalue())])
    assert np.array_equal(iters, g["iters"])
    assert relmax(E0, g["E0_final"]) < 1e-11
    assert relmax(jb, g["jbias"]) < 1e-8
    h = N // 2
    xe = g["xe_series"][-1] if "xe_series" in g.files else None
    if xe is not None:
        assert relmax(res["x0"][:h], xe) < 1e-11
        assert relmax(res["x0"][h:], g["xi_series"][-1]) < 1e-11
    assert relmax(res["phih"], g["phi_series"][-1]) < 1e-9


def test_pic_l_dd_functions_module(golden):
    import PIC_L_DD
    g = golden("dd_kernels")
    tag = "a"
    Ng = int(g[f"{tag}_Ng"]); dx = float(g[f"{tag}_dx"]); x = g[f"{tag}_x"]; F = g[f"{tag}_F"]
    q = g[f"{tag}_q"]; v = g[f"{tag}_v"]; active = g[f"{tag}_active"]
    p2c = float(g[f"{tag}_p2c"]); dt = float(g[f"{tag}_dt"]); N = len(x)
    assert np.array_equal(PIC_L_DD.interpolateField(F, x, Ng, dx), g[f"{tag}_interp"])
    assert PIC_L_DD.interpolateField(F, float(x[5]), Ng, dx) == g[f"{tag}_interp"][5]      # scalar call, like the reference
    assert relmax(PIC_L_DD.weightCurrents(x, q, v, p2c, Ng, N, dx, dt, active), g[f"{tag}_j"]) < 1e-13
    assert relmax(PIC_L_DD.weightDensities(x, q, p2c, Ng, N, dx, active), g[f"{tag}_rho"]) < 1e-13
    assert np.array_equal(PIC_L_DD.differentiateField(F, dx, Ng), g[f"{tag}_diff"])
    assert relmax(PIC_L_DD.integrateField(F, dx, Ng), g[f"{tag}_int"]) < 1e-13
    assert np.array_equal(PIC_L_DD.smoothField(F), g[f"{tag}_smooth"])


# ------------------------------------------------------------------------------- PIC_L
def test_pic_l_functions_module(golden):
    import PIC_L
    g = golden("l_kernels")
    Ng = int(g["Ng"]); dx = float(g["dx"]); x = g["x"]; v = g["v"]; q = g["q"]; m = g["m"]; p2c = float(g["p2c"]); E = g["E"]
    N = len(x); L = dx * (Ng - 1)
    assert np.array_equal(PIC_L.interpolateFieldPeriodic(E, x, Ng, dx), g["interp"])
    assert PIC_L.interpolateFieldPeriodic(E, float(x[3]), Ng, dx) == g["interp"][3]
    assert relmax(PIC_L.weightDensitiesPeriodic(x, q, p2c, Ng, N, dx), g["rho"]) < 1e-13
    assert relmax(PIC_L.weightCurrentsPeriodic(x, q, v, p2c, Ng, N, dx), g["j"]) < 1e-13
    phi = PIC_L.solvePoissonPeriodicElectronsNeutralized(dx, Ng, g["rho"], 1.0, 1e-3, 20, np.zeros(Ng + 1))
    assert relmax(phi - np.max(phi), g["phi"]) < 1e-9
    assert np.array_equal(PIC_L.differentiateFieldPeriodic(g["phi"], dx, Ng), g["dphi"])
    xo, vo = PIC_L.pushParticlesExplicit(x, v, q, m, N, Ng, 1e-9, dx, E)
    assert np.array_equal(xo, g["xout"]) and np.array_equal(vo, g["vout"])                  # bit-exact
    xb, vb = PIC_L.applyBoundaryConditionsPeriodic(xo, vo, m, N, L, dx, 1.0)
    assert np.array_equal(xb, g["xbc"])


def test_pic_l_and_pic_l_dd_undriven_functions_golden(golden):
    """The module functions no reference driver calls (PIC_L.py:48-60, 83-98, 146-206, 261-282 and
    PIC_L_DD.py:116-176), against the reference's own output on the same inputs."""
    import PIC_L
    import PIC_L_DD
    g = golden("stub_functions")
    Ng = int(g["Ng"]); dx = float(g["dx"]); L = dx * (Ng - 1)
    x, q, m, v, p2c = g["x"], g["q"], g["m"], g["v"], float(g["p2c"]); N = len(x)
    assert relmax(PIC_L.weightCurrents(x, q, v, p2c, Ng, N, dx), g["l_j"]) < 1e-13
    assert relmax(PIC_L.weightDensities(x, q, p2c, Ng, N, dx), g["l_rho"]) < 1e-13
    kBT = float(g["kBT"])
    # bounded Boltzmann-Newton: pinned after a fixed number of iterations (tol = 0), see the generator
    assert relmax(PIC_L.solvePoisson(dx, Ng, g["rho_b"], kBT, 0.0, 1, g["phi0_b"].copy()), g["l_phi_b1"]) < 1e-9
    assert relmax(PIC_L.solvePoisson(dx, Ng, g["rho_b"], kBT, 0.0, 3, g["phi0_b"].copy()), g["l_phi_b3"]) < 1e-8
    assert relmax(PIC_L_DD.solvePoisson(dx, Ng, g["rho_b"], kBT, 0.0, 3, g["phi0_b"].copy()), g["dd_phi_b3"]) < 1e-8
    assert relmax(PIC_L.solvePoissonPeriodic(dx, Ng, g["rho_p"], kBT, 1e-8, 20, np.zeros(Ng + 1)), g["l_phi_p"]) < 1e-9
    assert relmax(PIC_L_DD.solvePoissonPeriodic(dx, Ng, g["rho_p"][:Ng], kBT, 1e-8, 20, np.zeros(Ng)), g["dd_phi_p"]) < 1e-9
    xo, vo = PIC_L.pushParticlesImplicit(x, g["xh"], v, q, m, N, Ng, 1e-10, dx, g["Eh"])
    assert np.array_equal(xo, g["l_xi"]) and np.array_equal(vo, g["l_vi"])
    xb, vb = g["bc_x_in"].copy(), g["bc_v_in"].copy()
    np.random.seed(5)
    xb2, vb2 = PIC_L.applyBoundaryConditions(xb, vb, m, N, L, dx, kBT)
    assert xb2 is xb and np.array_equal(xb2, g["bc_x"]) and np.array_equal(vb2, g["bc_v"])
    assert np.random.uniform() == float(g["bc_next_uniform"])


def test_pygcpic_particle_ionisation_attempts_golden(golden):
    """Particle.attempt_first_ionization / attempt_nth_ionization (pygcpic.py:350-458) object by object:
    same charge states, same added_particles, same number of legacy-stream draws as the reference."""
    import pygcpic as G
    g = golden("stub_functions")
    grid = G.Grid(int(g["ion_ng"]), float(g["ion_L"]), 60. * 11600.)
    grid.n[:] = g["ion_n"]
    np.random.seed(9)
    for Z, cs, p2c, x, nth, cs_after, added in g["ion_rows"]:
        pt = G.Particle(G.mp * (10.81 if Z == 5 else 1.0), cs, p2c, 1.0, int(Z), grid=grid)
        pt.r[0] = x
        before = grid.added_particles
        with contextlib.redirect_stdout(io.StringIO()):
            (pt.attempt_nth_ionization if nth else pt.attempt_first_ionization)(2e-7, 60. * 11600., grid)
        assert float(pt.charge_state) == cs_after and grid.added_particles - before == added
    assert np.random.uniform() == float(g["ion_next_uniform"])
    with pytest.raises(UnboundLocalError):
        G.Particle(G.mp, 0, 1.0, 1.0, 1, grid=grid).attempt_nth_ionization(1e-7, 1e5, grid)


@pytest.mark.parametrize("sort_every", [None, 3])
def test_pic_l_main_module(golden, tmp_path, sort_every):
    """PIC_L.main(T, nplot) from the same seed as the golden reference run (N literal 6000); with the cell sort
    (what main does by itself from 2^17 particles on) the window kernel runs and the particles come back in
    the reference's order."""
    import PIC_L
    g = golden("l_main")
    N = int(g["N"]); T = int(g["T"])
    res = {}
    np.random.seed(1)
    with scratch_cwd(str(tmp_path)):
        EE = PIC_L.main(T, 1, N=N, result=res, sort_every=sort_every)
        EE_file = np.loadtxt("plots/E2.txt")
    assert relmax(EE, g["EE"]) < 1e-9
    assert np.array_equal(EE_file, np.array(EE))
    assert relmax(res["E_series"], g["E_series"]) < 1e-9
    if sort_every:
        ref = {}
        np.random.seed(1)
        with scratch_cwd(str(tmp_path)):
            PIC_L.main(T, 1, N=N, result=ref, sort_every=0)
        assert relmax(res["x"], ref["x"]) < 1e-12 and relmax(res["v"], ref["v"]) < 1e-10


# ------------------------------------------------------------------------------- pypic
def test_pypic_particle_push_module(golden):
    """pypic.particle_push_p(...) as the reference's implicit_pic calls it, three consecutive steps."""
    import pypic
    g = golden("pypic_push")
    N = int(g["N"]); Ng = int(g["Ng"]); L = float(g["L"]); dx = float(g["dx"]); dt = float(g["dt"])
    q = -np.ones(N) * pypic.e; m = np.ones(N) * pypic.me
    x, v, E, j = g["x0"], g["v0"], g["E0"], g["j0"]
    with contextlib.redirect_stdout(io.StringIO()):
        for s in range(3):
            x, v, E, j = pypic.particle_push_p(x, v, q, m, E, j, N, Ng, float(g["p2c"]), dx, dt, L, float(g["tol"]),
                                               int(g["maxiter"]))
            assert relmax(x, g[f"x_{s}"]) < 1e-12 and relmax(v, g[f"v_{s}"]) < 1e-12
            assert relmax(E, g[f"E_{s}"]) < 1e-11 and relmax(j, g[f"j_{s}"]) < 1e-11


def test_pypic_main_module_runs(tmp_path):
    """pypic.main(T, nplot) end to end (what run_pypic.py calls) on a reduced particle count:
    writes E2.txt / J.txt / parameters.out and conserves total energy to a few per cent."""
    import pypic
    res = {}
    np.random.seed(1)
    with scratch_cwd(str(tmp_path)):
        pypic.main(6, 10, N=40000, Ng=64, result=res)
        assert os.path.isfile("plots/E2.txt") and os.path.isfile("plots/J.txt") and os.path.isfile("plots/parameters.out")
        EE = np.loadtxt("plots/E2.txt")
    assert len(EE) == 6 and np.all(np.isfinite(EE)) and np.all(EE > 0)
    tot = res["EE"] + res["KE"]
    assert abs(tot[-1] - tot[0]) < 0.05 * abs(tot[0])
    # the same run on the sorted store (implicit_pic's own choice from 2^17 particles on): the window kernel runs,
    # particles and series come back in the reference's order
    srt = {}
    np.random.seed(1)
    with scratch_cwd(str(tmp_path)):
        pypic.main(6, 10, N=40000, Ng=64, result=srt, sort_every=2)
    assert relmax(srt["EE"], res["EE"]) < 1e-10 and relmax(srt["KE"], res["KE"]) < 1e-12
    assert relmax(srt["x0"], res["x0"]) < 1e-12 and relmax(srt["v0"], res["v0"]) < 1e-10
    assert relmax(srt["E0"], res["E0"]) < 1e-10
    assert res["x0"].min() >= 0.0 and res["x0"].max() <= 22.0 * np.sqrt(pypic.kb * 100.0 * 11600. * pypic.epsilon0 / pypic.e**2 / 1e5)


def test_launch_script_shape_runs_unchanged(tmp_path):
    """tools/drive.py runs a launcher with the statements of the reference's run_pypic_dd.py
    (`import PIC_L_DD as p; import convert as c; p.main_i(stop, skip)`) against the drop-in
    modules.  The launcher is written here with a short T (the reference's literals would
    take hours); the imageio-based GIF step is skipped when imageio is absent."""
    import subprocess
    launcher = os.path.join(str(tmp_path), "run_pypic_dd.py")
    with open(launcher, "w") as f:
        f.write("import PIC_L_DD as p\nimport convert as c\n\ndef main():\n\tstart = 0\n\tstop = 3\n\tskip = 10\n"
                "\tp.main_i(stop,skip)\n\nif __name__ == '__main__':\n\tmain()\n")
    os.makedirs(os.path.join(str(tmp_path), "plots"), exist_ok=True)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "drive.py"), launcher, "--seed", "1"],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("Iterations:") == 4
    assert os.path.isfile(os.path.join(str(tmp_path), "E0.txt"))


# ------------------------------------------------------------------------------- pygcpic
def test_pygcpic_doctest_kats():
    """The known-answer doctests the reference ships for the hot path,
    evaluated against the drop-in classes: every number comes out of a CUDA kernel."""
    import pygcpic as G
    e = G.e
    # interpolate_electric_field_dirichlet :336-342 and push_6D :469-476 (shared default E0 quirk)
    grid = G.Grid(100, 1.0, 1.0)
    grid.E[:] = 1.0
    p = G.Particle(1.0, 1.0, 1.0, 1.0, 1, grid=grid)
    p.interpolate_electric_field_dirichlet(grid)
    assert p.E[0] == 1.0
    p2 = G.Particle(1.0, 1 / e, 1.0, 1.0, 1)
    assert p2.E[0] == 1.0                              # the mutable default array is shared, as in the reference
    p2.push_6D(1.0)
    assert round(p2.r[3], 6) == 1.0
    G.Particle.__init__.__defaults__[1][:] = 0.0       # undo for the tests below
    # weight_particles_to_grid_boltzmann :852-866
    grid = G.Grid(101, 1.0, 1.0)
    pa = G.Particle(1.0, 1.0, 1.0, 1.0, 1, grid=grid); pa.r[0] = 0.0
    grid.weight_particles_to_grid_boltzmann([pa], 1.0)
    assert grid.n[0] == 100.0
    pa.r[0] = 1.0 - grid.dx / 2
    grid.weight_particles_to_grid_boltzmann([pa], 1.0)
    assert round(grid.n[-1], 6) == 50.0
    # differentiate_phi_to_E_dirichlet :922-930
    grid = G.Grid(6, 5.0, 1.0)
    grid.phi[:] = 1.0
    grid.differentiate_phi_to_E_dirichlet()
    assert np.all(np.abs(grid.E) < 1e-15)
    grid.phi[:] = np.linspace(0.0, 1.0, 6)
    grid.differentiate_phi_to_E_dirichlet()
    assert np.allclose(grid.E, -0.2, rtol=0, atol=1e-15)
    # solve_for_phi_dirichlet :992-996
    grid = G.Grid(5, 4.0, 1.0)
    grid.rho[:] = 1.0
    grid.solve_for_phi_dirichlet()
    assert np.allclose(grid.phi, [0.0, 1.5, 2.0, 1.5, 0.0], rtol=0, atol=1e-14)
    # neutral plasma -> phi == 0 for both Newton-Boltzmann solves :1013-1019, :1070-1076
    grid = G.Grid(5, 4.0, 1.0)
    grid.n0 = 1.0 / e
    grid.rho[:] = np.ones(5)
    grid.n[:] = np.ones(5) / e
    grid.solve_for_phi_dirichlet_boltzmann()
    assert np.array_equal(grid.phi, np.zeros(5))
    grid = G.Grid(5, 4.0, 1.0)
    grid.n0 = 1.0 / e * G.epsilon0
    grid.rho[:] = np.ones(5)
    grid.n[:] = np.ones(5) / e * G.epsilon0
    grid.solve_for_phi_dirichlet_neumann_boltzmann()
    assert np.array_equal(grid.phi, np.zeros(5))
    with contextlib.redirect_stdout(io.StringIO()):
        gdn = G.Grid(5, 4.0, 1.0, bc='dirichlet-neumann')
    assert gdn.A[-1, -1] == 3. and gdn.A[-1, -2] == -4. and gdn.A[-1, -3] == 1.
    # apply_BCs_periodic :655-663, apply_BCs_dirichlet :677-683
    grid = G.Grid(5, 1.0, 1.0)
    pb = G.Particle(1.0, 1.0, 1.0, 1.0, 1, grid=grid)
    pb.r[0] = grid.length * 1.5
    pb.apply_BCs_periodic(grid)
    assert pb.is_active() and pb.r[0] == grid.length * 0.5
    pb.r[0] = grid.length + 1.0
    pb.apply_BCs_dirichlet(grid)
    assert not pb.is_active() and pb.at_wall == 1
    # constructor errors :803-806
    with pytest.raises(ValueError):
        G.Grid(5, 1.0, 1.0, bc='periodic')
    with pytest.raises(TypeError):
        G.Grid(5, 1.0, 1.0, bc=3)


def test_pygcpic_particle_methods_golden(golden):
    """Particle.interpolate/push_6D/transform_6D_to_GC/push_GC/transform_GC_to_6D, one object per
    particle like the reference, against the survey's golden vectors and the batch golden."""
    import pygcpic as G
    g = golden("gc")
    B = g["B"]
    pt = G.Particle(G.mp, 1, 1.0, 1.0, 1, B0=B.copy(), E0=np.array([1000.0, 0.0, 0.0]))
    pt.r[:] = [1e-4, 0, 0, 1e4, 2e4, -3e4, 0]
    pt.push_6D(1e-10)
    assert relmax(pt.r, g["boris_one"]) < 1e-15
    grid = G.Grid(int(g["ng"]), float(g["Lg"]), 60. * 11600.)
    grid.E[:] = g["grid_E"]
    np.random.seed(17)
    for i in range(0, 60):
        pt = G.Particle(float(g["ms"][i]), int(g["cs"][i]), 1.0, 1.0, 1, B0=B.copy(), E0=g["Eshared"].copy())
        pt.r[:] = g["r0"][i]
        pt.interpolate_electric_field_dirichlet(grid)
        assert pt.E[0] == g["gather"][i]
        pt.push_6D(1e-10)
        assert relmax(pt.r, g["r_boris"][i]) < 1e-14
        if g["cs"][i] != 0:
            pt.transform_6D_to_GC()
            assert relmax(pt.r, g["r_gc"][i]) < 1e-13 and pt.mode == 1
            pt.push_GC(1e-10)
            assert relmax(pt.r, g["r_gc2"][i]) < 1e-12
            pt.transform_GC_to_6D()                      # draws a from the global stream (seeded 17 like the golden)
            assert relmax(pt.r, g["r_back"][i]) < 1e-12 and pt.mode == 0
        else:
            np.random.uniform(0.0, 1.0, 0)


def test_pygcpic_run_sheath_golden(golden):
    """pygcpic.run_sheath (device-resident loop of pic_bca_aps' particle phase) vs the reference's
    object loop: identical integer outcomes per step, fields and particles to tolerance."""
    import pygcpic as G
    g = golden("gc")
    Ld = float(g["drv_L"]); ngd = int(g["drv_ng"]); Nd = int(g["drv_N"]); dt = float(g["drv_dt"])
    p2c = float(g["drv_p2c"]); Ti = float(g["drv_Ti"]); Te = float(g["drv_Te"]); source_N = int(g["drv_source_N"])
    np.random.seed(int(g["drv_seed"]))
    host_grid = G.Grid(ngd, Ld, Te)
    parts = [G.Particle(G.mp, 1, p2c, Ti, Z=1, B0=g["B"].copy(), E0=np.zeros(3), grid=host_grid) for _ in range(Nd)]
    r = np.array([p.r for p in parts])
    assert np.array_equal(r, g["drv_r_init"])
    grid = G.GridDev(ngd, Ld, Te)
    st = G.ParticleStore.from_arrays(r, 1.0, G.mp, p2c, Z=1, B=g["B"])
    src = G.source_distribution_6D(host_grid, Ti, G.mp)
    out = G.run_sheath(grid, st, dt, 25, source_N, src, p2c, G.mp)
    assert np.array_equal(out["length"], g["drv_len"]) and np.array_equal(out["hits"], g["drv_hits"])
    assert np.array_equal(out["deleted"], g["drv_ndel"]) and np.array_equal(out["reactivated"], g["drv_nreact"])
    assert relmax(out["n0"], g["drv_n0"]) < 1e-9
    assert np.array_equal(st.flags_host()["active"], g["drv_active_final"])
    assert relmax(st.r_host(), g["drv_r_final"]) < 1e-6
    assert relmax(np.concatenate(out["ekin"]), g["drv_ekin"]) < 1e-5
    assert relmax(np.concatenate(out["angle"]), g["drv_ang"]) < 1e-5


def test_pygcpic_host_grid_n0_update_matches_device_grid():
    """Grid.weight_particles_to_grid_boltzmann (host attributes) vs GridDev on the second call,
    when phi != 0 enters the reference-density update (regression: a staging tensor freed before
    the launch made the kernel read the domain array as phi)."""
    import torch
    import pygcpic as G
    rs = np.random.RandomState(6)
    N, ng, Lg, Te, dt = 4000, 50, 2e-3, 7e5, 1e-10
    r = np.zeros((N, 7)); r[:, 0] = rs.uniform(0, Lg, N)
    st = G.ParticleStore.from_arrays(r, 1.0, G.mp, 2e9, Z=1)
    hg = G.Grid(ng, Lg, Te)
    dg = G.GridDev(ng, Lg, Te)
    phi = rs.uniform(0, 40, ng)
    for call in range(2):
        hg.weight_particles_to_grid_boltzmann(st, dt)
        dg.weight_particles_to_grid_boltzmann(st, dt)
        assert abs(hg.n0 - dg.n0) <= 1e-13 * abs(dg.n0)
        hg.phi = phi.copy(); dg.phi.copy_(torch.as_tensor(phi))
        hg.add_particles(2e9); dg.add_particles(2e9)
    # closed form of pygcpic.py:895-903 for the second call
    eta = np.exp(phi / Te / 11600.)
    assert np.isfinite(hg.n0) and hg.n0 > 0


@pytest.mark.parametrize("fused_min,event_loop", [(None, False), (0, False), (None, True)])
def test_pygcpic_run_sheath_with_ionisation_golden(golden, fused_min, event_loop):
    """N3: Monte-Carlo ionisation of neutral H and B(0..2), mid-domain exits of wall-born particles
    and the reactivate-or-delete rule coupled through the running source-ion count, against the
    reference's own objects driven through pic_bca_aps' particle loop (tests/golden/gc_ion.npz,
    oracle/make_golden.py::gen_gc_ion): identical integer outcomes per step, identical number of
    RNG draws consumed."""
    import pygcpic as G
    g = golden("gc_ion")
    Ld = float(g["L"]); ngd = int(g["ng"]); Nd = int(g["N"]); dt = float(g["dt"]); p2c = float(g["p2c"])
    Ti = float(g["Ti"]); Te = float(g["Te"]); source_N = int(g["source_N"]); B = g["B"]
    n_ion, n_h0, n_b = 1500, 300, 300
    p2c_n = p2c * 2e-4
    np.random.seed(41)
    host_grid = G.Grid(ngd, Ld, Te)
    kinds = np.array([0] * n_ion + [1] * n_h0 + [2] * n_b)
    np.random.shuffle(kinds)
    parts = []
    for kd in kinds:                       # same construction (and draw) order as the generator
        if kd == 0:
            p_ = G.Particle(G.mp, 1, p2c, Ti, Z=1, B0=B.copy(), E0=np.zeros(3), grid=host_grid)
        elif kd == 1:
            p_ = G.Particle(G.mp, 0, p2c_n, Ti, Z=1, B0=B.copy(), E0=np.zeros(3), grid=host_grid)
            p_.from_wall = int(np.random.uniform() < 0.5)
        else:
            p_ = G.Particle(10.81 * G.mp, int(np.random.randint(0, 3)), p2c_n, Ti, Z=5, B0=B.copy(), E0=np.zeros(3), grid=host_grid)
            p_.from_wall = int(np.random.uniform() < 0.5)
        parts.append(p_)
    r = np.array([p_.r for p_ in parts])
    assert np.array_equal(r, g["r_init"])
    assert np.array_equal([float(p_.charge_state) for p_ in parts], g["cs_init"])
    assert np.array_equal([p_.from_wall for p_ in parts], g["from_wall_init"])
    st = G.ParticleStore.from_arrays(r, g["cs_init"], g["m_init"], g["p2c_init"], Z=g["Z_init"],
                                     from_wall=g["from_wall_init"], B=B)
    if fused_min is not None:
        st.FUSED_MIN = fused_min
    grid = G.GridDev(ngd, Ld, Te)
    src = G.source_distribution_6D(host_grid, Ti, G.mp)
    # fused_min=0: the mixed-species fused kernel's ABI (its per-particle routine at this size); event_loop: the
    # global-event-list formulation of the decisions that sharded runs use
    out = G.run_sheath(grid, st, dt, 20, source_N, src, p2c, G.mp, ionize_Te=Te, event_loop=event_loop)
    assert np.array_equal(out["length"], g["h_length"]) and np.array_equal(out["hits"], g["h_hits"])
    assert np.array_equal(out["deleted"], g["h_ndel"]) and np.array_equal(out["reactivated"], g["h_nreact"])
    assert np.array_equal(out["ionised_h"], g["h_nion_h"]) and np.array_equal(out["ionised_b"], g["h_nion_b"])
    assert np.array_equal(out["midexit"], g["h_nexit"])
    assert relmax(out["n0"], g["h_n0"]) < 1e-9
    assert np.array_equal(st.charge_state[:st.N].cpu().numpy(), g["cs_final"])
    assert np.array_equal(st.Z[:st.N].cpu().numpy(), g["Z_final"])
    assert np.array_equal(st.flags_host()["active"], g["active_final"])
    assert relmax(st.r_host(), g["r_final"]) < 1e-6
    assert np.random.uniform() == float(g["next_uniform"])        # exactly as many draws as the reference


def test_pygcpic_run_sheath_ionisation_on_a_fused_eligible_store():
    """A species-uniform store large enough for the fused push kernel, WITH Monte-Carlo ionisation and
    wall-born particles: the fused kernel must not pre-deposit for the next step (charge states change,
    mid-domain exits deactivate particles after the push), so the run has to equal the same run on the
    per-particle kernels -- integer tallies exactly, n0 to round-off, same number of RNG draws."""
    import pygcpic as G
    N, ng, Lg = 3 * 16384 + 123, 150, 6e-3
    Te, Ti = 60. * 11600., 50. * 11600.
    B = np.array([2 * np.cos(86 * np.pi / 180), 2 * np.sin(86 * np.pi / 180), 0.0])
    rs = np.random.RandomState(8)
    r = np.zeros((N, 7)); r[:, 0] = rs.uniform(0.02 * Lg, 0.98 * Lg, N)
    r[:, 3:6] = rs.normal(0, np.sqrt(G.kb * Ti / G.mp), (N, 3))
    fw = (rs.uniform(size=N) < 0.5).astype(np.int8)
    p2c = 4e15 * Lg / N
    host_grid = G.Grid(ng, Lg, Te)

    def run(fused):
        np.random.seed(3)
        st = G.ParticleStore.from_arrays(r.copy(), np.zeros(N), np.full(N, G.mp), np.full(N, p2c), Z=np.ones(N, dtype=np.int32),
                                         from_wall=fw, B=B)
        if not fused:
            st.FUSED_MIN = 10**12
        assert (st.uniform() is not None) and (st.N >= G.ParticleStore.FUSED_MIN)
        grid = G.GridDev(ng, Lg, Te)
        src = G.source_distribution_6D(host_grid, Ti, G.mp)
        out = G.run_sheath(grid, st, 2e-9, 6, N // 2, src, p2c, G.mp, charge_state=1, Z=1, ionize_Te=Te)
        return out, st, np.random.uniform()
    a, sa, ua = run(True)
    b, sb, ub = run(False)
    assert ua == ub
    for k in ("length", "hits", "deleted", "reactivated", "ionised_h", "midexit"):
        assert np.array_equal(a[k], b[k]), k
    assert sum(a["ionised_h"]) > 0 and sum(a["midexit"]) > 0
    assert relmax(a["n0"], b["n0"]) < 1e-10
    assert np.array_equal(sa.charge_state[:sa.N].cpu().numpy(), sb.charge_state[:sb.N].cpu().numpy())
    assert relmax(sa.r_host(), sb.r_host()) < 1e-9
