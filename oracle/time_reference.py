"""TEST INFRASTRUCTURE ONLY -- times the reference's OWN hot path (through oracle/refshim.py) on the
authoring container's CPU, for the "reference itself" row of DESIGN.md section 7 ("CPU
baseline timing").  /root/reference does not travel to the GPU box, so bench.py's reference arm times
the C/OpenMP port instead; this script records how the actual Python reference compares.

    python -m oracle.time_reference      -> tests/golden/reference_timings.json
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refshim, np_oracle as O  # noqa: E402


def main():
    out = {"host": "authoring container", "cores": os.cpu_count()}
    # --- pypic.particle_push_p at the shipped default size (numba kernels for gather/deposit)
    p = refshim.load("pypic")
    import numba
    N, Ng = 1_000_000, 200
    L = 5170.094; dx = L / Ng; dt = 1e-5
    rs = np.random.RandomState(1)
    x0 = rs.uniform(0, L, N); v0 = rs.normal(0, 4.2e6, N)
    q = -np.ones(N) * O.e; m = np.ones(N) * O.me
    E0 = rs.normal(0, 1e-3, Ng); j0 = np.zeros(Ng)
    with contextlib.redirect_stdout(io.StringIO()):
        p.particle_push_p(x0, v0, q, m, E0, j0, N, Ng, 5170, dx, dt, L, 1e-3, 20)      # JIT warm-up
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            p.particle_push_p(x0, v0, q, m, E0, j0, N, Ng, 5170, dx, dt, L, 1e-3, 20)
        t = (time.perf_counter() - t0) / reps
    out["pypic.particle_push_p"] = dict(N=N, Ng=Ng, s_per_step=t, particle_steps_per_s=N / t,
                                        numba_threads=numba.get_num_threads())
    # --- PIC_L_DD: the Picard loop body at the shipped default size, timed through its own functions
    d = refshim.load("PIC_L_DD")
    N, Ng = 40000, 51
    dxs = 1e-5; Ls = dxs * (Ng - 1); dts = 1e-12
    x0 = rs.uniform(0, Ls, N)
    mm = np.concatenate([np.full(N // 2, O.me), np.full(N // 2, O.mp)])
    qq = np.concatenate([np.full(N // 2, -O.e), np.full(N // 2, O.e)])
    u0 = rs.normal(0, 1, N) * np.sqrt(O.kb * 116000. / mm)
    Es = rs.normal(0, 1e4, Ng)
    act = np.ones(N)
    t0 = time.perf_counter()
    # one Picard iteration as written in PIC_L_DD.py:470-513 (gather loop, push, two deposits)
    Ei = np.array([d.interpolateField(Es, x0[i], Ng, dxs) for i in range(N)])
    x1 = x0 + dts * u0 + dts * dts * (qq / mm) * Ei * 0.5
    u1 = u0 + dts * (qq / mm) * Ei
    xh = (x0 + x1) * 0.5; uh = (u0 + u1) * 0.5
    d.weightCurrents(np.clip(xh, 0, Ls * (1 - 1e-12)), qq, uh, 1.25e11, Ng, N, dxs, dts, act)
    d.weightCurrents(np.clip(x1, 0, Ls * (1 - 1e-12)), qq, u1, 1.25e11, Ng, N, dxs, dts, act)
    t_iter = time.perf_counter() - t0
    out["PIC_L_DD.picard_iteration"] = dict(N=N, Ng=Ng, s_per_iteration=t_iter, iterations_per_step=5,
                                            particle_steps_per_s=N / (5 * t_iter))
    # --- pygcpic: gather + push_6D + BC per particle object
    g = refshim.load("pygcpic")
    grid = g.Grid(150, 1.8e-3, 60 * 11600.)
    grid.E[:] = rs.normal(0, 1e4, 150)
    B = np.array([2 * np.cos(1.5), 2 * np.sin(1.5), 0.])
    parts = [g.Particle(g.mp, 1, 1e9, 50 * 11600., Z=1, B0=B.copy(), E0=np.zeros(3), grid=grid) for _ in range(10000)]
    t0 = time.perf_counter()
    for pt in parts:
        pt.interpolate_electric_field_dirichlet(grid); pt.push_6D(1e-10); pt.apply_BCs_dirichlet(grid)
    grid.weight_particles_to_grid_boltzmann(parts, 1e-10)
    t = time.perf_counter() - t0
    out["pygcpic.push_6D+weight"] = dict(N=len(parts), s_per_step=t, particle_steps_per_s=len(parts) / t)
    path = os.path.join(ROOT, "tests", "golden", "reference_timings.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
