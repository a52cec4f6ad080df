"""Multi-GPU parity on real devices (skipped on a single-GPU box): tools/mgpu_check.py under
torchrun with 2 ranks -- particle-decomposed sheath vs the same global state on one GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sheath_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py"), "200000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    out = [l for l in lines if "iters_sharded" in l][-1]
    assert out["ok"], out
    assert out["iters_sharded"] == out["iters_single"]
    tracked = [l["tracked"] for l in lines if "tracked" in l]
    assert len(tracked) == 2 and all(t["ok"] for t in tracked), tracked
    det = [l for l in lines if "det" in l][-1]["det"]
    assert det["ok"], det
    per = [l for l in lines if "periodic" in l][-1]["periodic"]
    assert per["ok"], per
    bor = [l for l in lines if "boris" in l][-1]["boris"]
    assert bor["ok"], bor


def test_two_rank_peer_memory_reduction_matches_nccl():
    """SheathSim(reduce="p2p"): the field kernel sums the ranks' accumulators over NVLink peer
    memory (CUDA IPC buffers, flag handshake) instead of an NCCL all-reduce (tools/p2p_check.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "tools", "p2p_check.py"), "400000", "5"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"], out


def test_two_rank_pygcpic_run_sheath_matches_single_rank():
    """pygcpic.run_sheath with the particle list sharded over 2 ranks (all-reduced deposits, decisions over the
    global event list, RNG / source generator consumed as by one process) vs one rank: tools/gc_sharded_check.py."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29536", os.path.join(ROOT, "tools", "gc_sharded_check.py"), "60000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"], out
    assert sum(out["uniform_fused"]["reactivated"]) > 0 and sum(out["mixed_ionising"]["ionised_h"]) > 0


def test_two_rank_slab_decomposition_matches_single_rank():
    """Spatial (slab) decomposition: halo exchange + particle migration + routed re-injection on 2
    ranks vs the same global particles on one rank (tools/slab_check.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tools", "slab_check.py"), "200000", "513"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"], out
    assert out["stat"]["migrated"] > 0 and out["stat"]["exported"] > 0


def test_slab_sim_single_rank_conserves_and_matches_particle_path():
    """World size 1 (runs on the single-GPU box): the slab driver (per-species blocks, sort with
    headroom offset, ordinal-keyed re-injection) against SheathSim on the first step, then
    several steps with sorts: particle count conserved, every slot alive after re-injection."""
    import numpy as np
    from oracle import np_oracle as O
    from pypic_b200.dist import Comm
    from pypic_b200.sheath import SheathSim
    from pypic_b200.spatial import SlabSheathSim
    N, Ng = 120000, 257
    dx, dt = 1e-5, 1e-12
    L = dx * (Ng - 1)
    kT = O.kb * 116000.
    rs = np.random.RandomState(4)
    h = N // 2
    x0 = rs.uniform(0, L, N)
    u0 = np.concatenate([rs.normal(0, np.sqrt(kT / O.me), h), rs.normal(0, np.sqrt(kT / O.mp), N - h)])
    E0 = rs.normal(0, 1e4, Ng)
    p2c = L * 1e19 / N
    a = SlabSheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), comm=Comm(enabled=False), sort_every=2, seed=3)
    a.upload(x0, u0, E0)
    b = SheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), carry_vw=False, comm=Comm(enabled=False))
    b.upload(x0, u0, E0=E0)
    ka, _ = a.step(); kb_, _ = b.picard()
    assert ka == kb_
    Ea, Eb = a.E0.cpu().numpy(), b.E0.cpu().numpy()
    assert np.max(np.abs(Ea - Eb)) <= 1e-12 * np.max(np.abs(Eb))
    dead_seen = 0
    for _ in range(5):
        dead_seen += sum(int((blk.active[:blk.n] != 1).sum().item()) for blk in a.blocks)
        k, r = a.step()
        assert 1 <= k <= 20 and a.local_particles() == N
    a.check()
    assert dead_seen > 0
    for blk in a.blocks:
        x = blk.x0[:blk.n].cpu().numpy()
        assert blk.off % 2 == 0 and np.isfinite(x).all()


class _RankView:
    """rank / world of an emulated rank (no process group: the test moves the messages itself)."""

    def __init__(self, rank, world):
        self.rank, self.world, self.group, self.enabled = rank, world, None, False


@pytest.mark.parametrize("world,shift", [(2, 0), (3, 7), (4, -12)])
def test_slab_distributed_field_update_on_emulated_ranks(world, shift):
    """The distributed field update of the slab decomposition (csrc/slab_kernels.cu) with several ranks EMULATED on
    one GPU: every rank object runs its own particle and field kernels on its own buffers, the test performs the
    two all-gathers of an iteration by concatenating the ranks' messages.  Against the same global particles on one
    rank with the replicated field update (the round-1 path, itself checked against SheathSim): identical Picard
    iteration counts and absorbed counts, E / j1 / particles to round-off, and bit-identical E on the nodes two
    neighbours both compute.  `shift` assigns particles to ranks as if they had drifted that many cells since the
    last migration, so the guard bands carry real current."""
    import numpy as np
    import torch
    from oracle import np_oracle as O
    from pypic_b200.dist import Comm
    from pypic_b200.spatial import SlabSheathSim, slab_bounds
    N, Ng, G = 150000, 513, 16
    dx, dt = 1e-5, 1e-12
    L = dx * (Ng - 1)
    kT = O.kb * 116000.
    rs = np.random.RandomState(11 + world)
    h = N // 2
    x0 = rs.uniform(0, L, N)
    u0 = np.concatenate([rs.normal(0, np.sqrt(kT / O.me), h), rs.normal(0, np.sqrt(kT / O.mp), N - h)])
    E0 = rs.normal(0, 1e4, Ng)
    p2c = L * 1e19 / N
    one = SlabSheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), comm=Comm(enabled=False), sort_every=0, field="replicated")
    one.upload(x0, u0, E0)
    k1, r1 = one.picard()
    assert k1 >= 3
    # the emulated ranks: ownership by the cell `shift` cells away from the particle's
    cb = slab_bounds(Ng, world)
    cell = np.clip(np.floor(x0 / dx).astype(np.int64) + shift, 0, Ng - 2)
    owner = np.searchsorted(np.asarray(cb[1:-1]), cell, side="right")
    sims = []
    for r in range(world):
        sim = SlabSheathSim(N, Ng, dx, dt, p2c, kBT=(kT, kT), comm=_RankView(r, world), sort_every=0, guard=G)
        for sp, sl in ((0, slice(0, h)), (1, slice(h, N))):
            keep = owner[sl] == r
            blk = sim.blocks[sp]
            blk.n = int(keep.sum()); blk.cur, blk.off = 0, 0
            blk.X[0][:blk.n].copy_(torch.as_tensor(np.ascontiguousarray(x0[sl][keep])))
            blk.U[0][:blk.n].copy_(torch.as_tensor(np.ascontiguousarray(u0[sl][keep])))
            blk.active.fill_(1)
        sim.E0.copy_(torch.as_tensor(E0))
        sims.append(sim)
    for sim in sims:
        sim.begin_step()
    k = 0
    while True:
        for sim in sims:
            sim.iter_particles(k)
        gath = torch.cat([sim.msg for sim in sims])
        for sim in sims:
            sim.gath.copy_(gath); sim.iter_field()
        gath2 = torch.cat([sim.part for sim in sims])
        for sim in sims:
            sim.gath2.copy_(gath2); sim.iter_finish()
        outs = [sim.outcome() for sim in sims]
        assert all(o == outs[0] for o in outs)          # every rank holds the same residual history
        k, hist = outs[0]
        if not (hist[-1] > sims[0].tol and k < sims[0].maxiter):
            break
    for sim in sims:
        assert int(sim.ctl.item()) == 1
        sim.end_step(k); sim.check()
    assert k == k1
    E1 = one.E0.cpu().numpy(); j1 = one.j0.cpu().numpy()
    scale, jscale = np.max(np.abs(E1)), np.max(np.abs(j1))
    # residual history: the last residuals are differences at round-off level of E, hence the absolute term
    for ra, rb in zip(hist, one.outcome()[1]):
        assert abs(ra - rb) <= 1e-9 * rb + 1e-14 * scale * np.sqrt(Ng)
    for r, sim in enumerate(sims):
        Er = sim.E0.cpu().numpy(); jr = sim.j0.cpu().numpy()
        b = slice(sim.b0, sim.b1)                        # own nodes + guard nodes
        assert np.max(np.abs(Er[b] - E1[b])) <= 1e-13 * scale
        assert np.max(np.abs(jr[b] - j1[b])) <= 1e-12 * jscale
        if r + 1 < world:                                # shared nodes: the neighbours agree bit for bit
            nb = sims[r + 1]
            sh = slice(nb.b0, sim.b1)
            assert np.array_equal(Er[sh], nb.E0.cpu().numpy()[sh])
        assert np.array_equal(sim.wall_cum.cpu().numpy(), one.wall_cum.cpu().numpy())
    assert abs(float(sims[0].stats[2].item()) - float(one.stats[2].item())) <= 1e-12 * float(one.stats[2].item())     # field energy
    assert abs(float(sims[0].stats[1].item()) - float(one.stats[1].item())) <= 1e-9 * jscale                         # mean j1
    for sp in range(2):
        xa = np.sort(np.concatenate([s_.blocks[sp].x0[:s_.blocks[sp].n].cpu().numpy() for s_ in sims]))
        xb = np.sort(one.blocks[sp].x0[:one.blocks[sp].n].cpu().numpy())
        assert np.max(np.abs(xa - xb)) <= 1e-13 * L
        da = sum(int((s_.blocks[sp].active[:s_.blocks[sp].n] != 1).sum().item()) for s_ in sims)
        db = int((one.blocks[sp].active[:one.blocks[sp].n] != 1).sum().item())
        assert da == db
