#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the particle-decomposed
sheath (contiguous shards, one fp64 all-reduce of [jh|j1|counts] per Picard iteration, host
MT19937 draw service) against the SAME global state advanced on one GPU without a process
group.  Iteration counts and absorbed tallies must be identical, fields/particles to
round-off.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200.dist import Comm  # noqa: E402
from pypic_b200.rng import LegacyDraws  # noqa: E402
from pypic_b200.sheath import SheathSim  # noqa: E402

KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27


rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400000
    Ng = 257; dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1); steps = 6
    kT = KB * 116000.
    rs = np.random.RandomState(3)
    h = N // 2
    x0 = rs.uniform(0, L, N)
    u0 = np.concatenate([rs.normal(0, np.sqrt(kT / ME), h), rs.normal(0, np.sqrt(kT / MP), N - h)])
    E0 = rs.normal(0, 1e4, Ng)
    p2c = L * 1e19 / N

    def run(comm, deposit="window"):
        np.random.seed(1)
        sim = SheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), carry_vw=False, comm=comm, device=dev, rng="host",
                        draws=LegacyDraws(), deposit=deposit)
        sim.upload(x0, u0, E0=E0)
        its, dead = [], []
        for _ in range(steps):
            nd = sim.reinject()
            k, r = sim.picard(); sim.t += 1
            its.append(k); dead.append(int(nd or 0))
        sim.check()
        return sim, its, dead
    sharded, its_s, dead_s = run(Comm())
    out = sharded.download()
    # gather the shards on rank 0
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, dict(x0=out["x0"], u0=out["u0"], active=out["active"], start=sharded.start))
    else:
        parts = [dict(x0=out["x0"], u0=out["u0"], active=out["active"], start=0)]
    if rank == 0:
        single, its_1, dead_1 = run(Comm(enabled=False))
        ref = single.download()
        xs = np.concatenate([p["x0"] for p in parts]); us = np.concatenate([p["u0"] for p in parts])
        act = np.concatenate([p["active"] for p in parts])
        res = dict(world=world, N=N, iters_sharded=its_s, iters_single=its_1, dead_local_rank0=dead_s, dead_single=dead_1,
                   E_rel=rel(out["E0"], ref["E0"]), x_rel=rel(xs, ref["x0"]), u_rel=rel(us, ref["u0"]),
                   flags_equal=bool(np.array_equal(act, ref["active"])))
        res["ok"] = bool(its_s == its_1 and res["flags_equal"] and res["E_rel"] < 1e-10 and res["x_rel"] < 1e-11)
        print(json.dumps(res))
    # ---- tracked order, sharded (the path PIC_L_DD.main_i takes): every rank sorts its shard by cell with the
    # original-index payload, the host draws follow the GLOBAL original-index order (thermostat with gamma != 0
    # included: every rank walks the whole legacy stream and applies the hits inside its range), carried v,w
    v0 = rs.normal(0, 1e5, N); w0 = rs.normal(0, 1e5, N)

    def run_tracked(comm, gamma):
        np.random.seed(5)
        sim = SheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), carry_vw=True, comm=comm, device=dev, rng="host",
                        draws=LegacyDraws(), sort_every=2, gamma=gamma, vion_after=1)
        sim.upload(x0, u0, v0, w0, E0=E0)
        its = [sim.step()[0] for _ in range(steps)]
        sim.check()
        vion = sim.collect_vionout()
        return sim, its, vion, np.random.uniform()
    for gamma in (0.0, 0.01):
        ts, its_t, vion_t, nxt_t = run_tracked(Comm(), gamma)
        ot = ts.download()
        keys = ("x0", "u0", "v0", "w0", "active")
        if world > 1:
            pt = [None] * world
            dist.all_gather_object(pt, {k_: ot[k_] for k_ in keys})
        else:
            pt = [{k_: ot[k_] for k_ in keys}]
        if rank == 0:
            t1, its_t1, vion_1, nxt_1 = run_tracked(Comm(enabled=False), gamma)
            o1 = t1.download()
            cat = lambda k_: np.concatenate([p_[k_] for p_ in pt])
            rt = dict(gamma=gamma, iters_sharded=its_t, iters_single=its_t1, sorts=ts._sorts, E_rel=rel(ot["E0"], o1["E0"]),
                      x_rel=rel(cat("x0"), o1["x0"]), u_rel=rel(cat("u0"), o1["u0"]), v_rel=rel(cat("v0"), o1["v0"]),
                      w_rel=rel(cat("w0"), o1["w0"]), flags_equal=bool(np.array_equal(cat("active"), o1["active"])),
                      stream_position_equal=bool(nxt_t == nxt_1), vionout_len=[len(vion_t), len(vion_1)],
                      vionout_rel=(rel(np.array(vion_t), np.array(vion_1)) if len(vion_t) == len(vion_1) and len(vion_1) else None))
            rt["ok"] = bool(its_t == its_t1 and rt["flags_equal"] and rt["stream_position_equal"] and rt["E_rel"] < 1e-10
                            and rt["x_rel"] < 1e-11 and rt["u_rel"] < 1e-10 and rt["v_rel"] < 1e-12 and rt["w_rel"] < 1e-12
                            and len(vion_t) == len(vion_1) and (rt["vionout_rel"] is None or rt["vionout_rel"] < 1e-10))
            print(json.dumps({"tracked": rt}))
    # ---- reproducible build, sharded: the currents are all-reduced as int64 fixed-point words, so two
    # runs give bit-identical fields, every rank holds the same bits, and the result matches the
    # single-GPU reproducible run to round-off (the window partial sums differ with the sharding)
    d1, its_d1, _ = run(Comm(), "window-det")
    d2, its_d2, _ = run(Comm(), "window-det")
    same_runs = bool(torch.equal(d1.E0, d2.E0) and torch.equal(d1.j0, d2.j0) and torch.equal(d1.x0, d2.x0)
                     and torch.equal(d1.u0, d2.u0) and torch.equal(d1.active, d2.active))
    Eb = d1.E0.cpu().numpy().tobytes()
    if world > 1:
        allE = [None] * world; allsame = [None] * world
        dist.all_gather_object(allE, Eb); dist.all_gather_object(allsame, same_runs)
    else:
        allE, allsame = [Eb], [same_runs]
    if rank == 0:
        s1, its_s1, _ = run(Comm(enabled=False), "window-det")
        det = dict(world=world, runs_bit_identical=bool(all(allsame)), ranks_bit_identical=bool(all(e == allE[0] for e in allE)),
                   iters_sharded=its_d1, iters_single=its_s1, E_rel=rel(d1.E0.cpu().numpy(), s1.E0.cpu().numpy()),
                   E_rel_vs_default=rel(d1.E0.cpu().numpy(), out["E0"]))
        det["ok"] = bool(det["runs_bit_identical"] and det["ranks_bit_identical"] and its_d1 == its_d2 == its_s1
                         and det["E_rel"] < 1e-10 and det["E_rel_vs_default"] < 1e-10)
        print(json.dumps({"det": det}))
    # ---- the two periodic codes: rho / [jh|j1] all-reduced per deposit, field phase replicated
    from pypic_b200.periodic import ExplicitSim, PeriodicImplicitSim
    Np, Ngp = N, 256
    Lp = dx * Ngp
    xp = rs.uniform(0, Lp, Np); vp = rs.normal(0, np.sqrt(kT / ME), Np)
    Ep = rs.normal(0, 1e4, Ngp)

    def run_pypic(comm):
        sim = PeriodicImplicitSim(Np, Ngp, dx, dt, Lp, Lp * 1e19 / Np, tol=1e-3, maxiter=20, comm=comm, device=dev)
        sim.upload(xp, vp, Ep)
        its = [sim.push()[0] for _ in range(4)]
        sim.check()
        return sim, its

    def run_explicit(comm):
        Le = dx * (Ngp - 1)
        sim = ExplicitSim(Np, Ngp, dx, dt * 50, (Le + dx) * 1e16 / Np, q=(-1.602e-19, 1.602e-19), m=(ME, MP), n_split=Np // 2,
                          comm=comm, device=dev)
        sim.upload(xp, np.concatenate([vp[:Np // 2], vp[Np // 2:] * np.sqrt(ME / MP)]))
        for _ in range(4):
            sim.step()
        sim.field_solve()
        sim.check()
        return sim
    sp, its_p = run_pypic(Comm())
    se = run_explicit(Comm())
    op, oe = sp.download(), se.download()
    if world > 1:
        parts2 = [None] * world
        dist.all_gather_object(parts2, dict(px=op["x0"], pv=op["v0"], ex=oe["x"], ev=oe["v"]))
    else:
        parts2 = [dict(px=op["x0"], pv=op["v0"], ex=oe["x"], ev=oe["v"])]
    if rank == 0:
        s1, its_1p = run_pypic(Comm(enabled=False)); e1 = run_explicit(Comm(enabled=False))
        r1, r2 = s1.download(), e1.download()
        cat = lambda k: np.concatenate([p[k] for p in parts2])
        res2 = dict(pypic_iters_sharded=its_p, pypic_iters_single=its_1p, pypic_E_rel=rel(op["E0"], r1["E0"]),
                    pypic_x_rel=rel(cat("px"), r1["x0"]), pypic_v_rel=rel(cat("pv"), r1["v0"]),
                    explicit_E_rel=rel(oe["E"], r2["E"]), explicit_x_rel=rel(cat("ex"), r2["x"]),
                    explicit_v_rel=rel(cat("ev"), r2["v"]))
        res2["ok"] = bool(its_p == its_1p and res2["pypic_E_rel"] < 1e-10 and res2["pypic_x_rel"] < 1e-11 and
                          res2["explicit_E_rel"] < 1e-9 and res2["explicit_x_rel"] < 1e-11)
        print(json.dumps(dict(periodic=res2)))
    # ---- pygcpic Boris path: particles sharded, deposited density all-reduced, field solve replicated
    from pypic_b200.gcstore import GridDev, ParticleStore
    from pypic_b200.dist import shard_range
    Nb, ngb, Lb, Teb = 300000, 257, 2.5e-3, 6e5
    rb = np.zeros((Nb, 7)); rb[:, 0] = np.sort(rs.uniform(0.02 * Lb, 0.98 * Lb, Nb)); rb[:, 3:6] = rs.normal(0, 7e4, (Nb, 3))
    Bv = (2 * np.cos(1.5), 2 * np.sin(1.5), 0.)

    def run_boris(comm, lo, hi):
        grid = GridDev(ngb, Lb, Teb, device=dev, comm=comm)
        st = ParticleStore.from_arrays(rb[lo:hi], 1.0, MP, Lb * 1e19 / Nb, Z=1, B=Bv, device=dev)
        st.FUSED_MIN = 0
        grid.weight_particles_to_grid_boltzmann(st, 1e-10)
        hits = 0
        for _ in range(4):
            if grid.have_fused_n:
                grid.finish_fused_deposit(1.0, 1e-10)
            grid.smooth_rho(); grid.reset_added_particles()
            grid.solve_for_phi_dirichlet_boltzmann(); grid.differentiate_phi_to_E_dirichlet()
            hits += st.push_6D(1e-10, grid, deposit=True)
        grid.finish_fused_deposit(1.0, 1e-10)
        st.check(); grid.check()
        return grid, st, hits
    lo, hi = shard_range(Nb, rank, world)
    gs, ss, hs = run_boris(Comm() if world > 1 else None, lo, hi)
    rloc = ss.r_host()
    if world > 1:
        pr = [None] * world
        dist.all_gather_object(pr, (rloc, hs))
    else:
        pr = [(rloc, hs)]
    if rank == 0:
        g1, s1, h1 = run_boris(None, 0, Nb)
        res3 = dict(n_rel=rel(gs.n.cpu().numpy(), g1.n.cpu().numpy()), phi_rel=rel(gs.phi.cpu().numpy(), g1.phi.cpu().numpy()),
                    n0_rel=abs(gs.n0 - g1.n0) / abs(g1.n0), r_rel=rel(np.concatenate([p[0] for p in pr]), s1.r_host()),
                    hits=[int(sum(p[1] for p in pr)), int(h1)])
        res3["ok"] = bool(res3["n_rel"] < 1e-12 and res3["phi_rel"] < 1e-9 and res3["r_rel"] < 1e-12 and res3["hits"][0] == res3["hits"][1])
        print(json.dumps(dict(boris=res3)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
