"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: contiguous particle shards, the
single all-reduce of the grid accumulators per deposit, and the sharded re-injection draw
service that keeps the legacy MT19937 stream identical to the reference's single process."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pypic_b200.dist import Comm, local_split, shard_range
from pypic_b200.rng import LegacyDraws, sheath_step_draws


def test_shard_range_partitions_in_order():
    for N in (0, 1, 7, 40000, 200000001):
        for world in (1, 2, 3, 8):
            edges = [shard_range(N, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == N
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]                       # contiguous, rank order == index order
            sizes = [e[1] - e[0] for e in edges]
            assert max(sizes) - min(sizes) <= 1


def test_local_split_follows_species_boundary():
    N, world = 1001, 4
    ns = N // 2
    tot0 = 0
    for r in range(world):
        a, b = shard_range(N, r, world)
        s = local_split(ns, a, b)
        assert 0 <= s <= b - a
        tot0 += s
        # every local index below s is species 1 globally, every one at or above is species 2
        assert all(a + i < ns for i in range(s)) and all(a + i >= ns for i in range(s, b - a))
    assert tot0 == ns


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = Comm()
        assert comm.rank == rank and comm.world == world
        # (1) one fp64 all-reduce of [jh | j1 | absorbed counts]: identical bits on every rank
        Ng = 33
        rs = np.random.RandomState(100 + rank)
        acc = torch.tensor(np.concatenate([rs.normal(size=2 * Ng), rs.randint(0, 50, 4).astype(np.float64)]))
        mine = acc.clone()
        comm.allreduce_sum(acc)
        gathered = [torch.zeros_like(acc) for _ in range(world)]
        dist.all_gather(gathered, acc)
        assert all(torch.equal(g, acc) for g in gathered)
        # (2) sharded re-injection draws
        N, L = 4000, 5e-4
        ns = N // 2
        dead_global = np.sort(np.random.RandomState(7).choice(N, 137, replace=False))
        a, b = shard_range(N, rank, world)
        dead_local = dead_global[(dead_global >= a) & (dead_global < b)] - a
        counts = comm.allgather_int(len(dead_local))
        sig = (1.3e6, 3.0e4)
        sigma_local = np.where(dead_local + a >= ns, sig[1], sig[0])
        draws = LegacyDraws(np.random.RandomState(1))
        xd, ud, vd, wd = sheath_step_draws(draws, counts, rank, N, sigma_local, L)
        after = draws.rng.uniform()               # stream position after the step must agree across ranks
        q.put((rank, mine.numpy(), acc.numpy(), dead_local + a, xd, ud, vd, wd, after, comm.max_float(float(rank))))
        comm.barrier()
    finally:
        dist.destroy_process_group()


def test_world2_gloo_allreduce_and_draw_service():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # all-reduce == sum of the per-rank accumulators, counts stay exact integers
    total = res[0][1] + res[1][1]
    assert np.array_equal(res[0][2], total) and np.array_equal(res[1][2], total)
    assert np.array_equal(total[-4:], np.round(total[-4:]))
    assert res[0][9] == 1.0 and res[1][9] == 1.0
    # draws: concatenating the ranks reproduces the reference's single sequential stream
    N, L, ns = 4000, 5e-4, 2000
    dead = np.concatenate([res[0][3], res[1][3]])
    assert np.array_equal(dead, np.sort(np.random.RandomState(7).choice(N, 137, replace=False)))
    rs = np.random.RandomState(1)
    for _ in range(N - len(dead)):               # thermostat: one uniform per active particle (PIC_L_DD.py:421)
        rs.uniform(0.0, 1.0)
    ref = []
    for i in dead:                               # PIC_L_DD.py:431-448
        s = 3.0e4 if i >= ns else 1.3e6
        ref.append((rs.uniform(0.0, L), rs.normal(0.0, s), rs.normal(0.0, s), rs.normal(0.0, s)))
    ref = np.array(ref)
    got = np.stack([np.concatenate([res[0][k], res[1][k]]) for k in (4, 5, 6, 7)], 1)
    assert np.array_equal(got, ref)
    nxt = rs.uniform()
    assert res[0][8] == nxt and res[1][8] == nxt


def _worker_det(rank, world, port, q):
    """Reproducible build, sharded: SheathSim._allreduce_acc on the layout
    [jh | j1 | 4 counts] fp64 + [hi(2Ng) | lo(2Ng)] int64 (gloo stands in for NCCL)."""
    import types
    from pypic_b200 import fixedpoint as FP
    from pypic_b200.sheath import SheathSim
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = Comm()
        Ng, n = 17, 6000
        e, p2c, dx = 1.602e-19, 2.048e9, 1e-5
        s = FP.scale_exponent((-e, e), p2c, dx)
        rs = np.random.RandomState(11)                     # the same global deposits on every rank
        node = rs.randint(0, 2 * Ng, n)
        v = rs.normal(0, 20.0, n)
        out = []
        for part in ("interleaved", "halves"):
            mine = (np.arange(n) % world == rank) if part == "interleaved" else (np.arange(n) * world // n == rank)
            hi, lo = FP.split(v[mine], s)
            H = np.zeros(2 * Ng, dtype=np.int64); Lo = np.zeros(2 * Ng, dtype=np.int64)
            np.add.at(H, node[mine], hi); np.add.at(Lo, node[mine], lo)
            acc = torch.zeros(2 * Ng + 4 + 4 * Ng, dtype=torch.float64)
            acc[2 * Ng:2 * Ng + 4] = torch.tensor([1.0 + rank, 2.0, 0.0, 5.0 * rank])
            acc[2 * Ng + 4:].view(torch.int64).copy_(torch.as_tensor(np.concatenate([H, Lo])))
            sim = types.SimpleNamespace(det=True, comm=comm, Ng=Ng, acc=acc, p2p=None)
            SheathSim._allreduce_acc(sim)
            out.append(acc.numpy().copy())
        q.put((rank, out, s))
        comm.barrier()
    finally:
        dist.destroy_process_group()


def test_world2_gloo_reproducible_build_allreduce_is_partition_independent():
    from pypic_b200 import fixedpoint as FP
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_det, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Ng, n = 17, 6000
    s = res[0][2]
    # both ranks and both partitions of the deposits end with the same bits
    ref = res[0][1][0]
    for r in res:
        for acc in r[1]:
            assert acc.tobytes() == ref.tobytes()
    assert np.array_equal(ref[:2 * Ng], np.zeros(2 * Ng))                       # the fp64 current slots are not used
    assert np.array_equal(ref[2 * Ng:2 * Ng + 4], [3.0, 4.0, 0.0, 5.0])         # counts: exact fp64 integers
    words = ref[2 * Ng + 4:].view(np.int64)
    rs = np.random.RandomState(11)
    node = rs.randint(0, 2 * Ng, n); v = rs.normal(0, 20.0, n)
    import math
    for g in range(2 * Ng):
        exact = math.fsum(v[node == g].tolist())
        got = FP.merge(int(words[g]), int(words[2 * Ng + g]), s)
        assert abs(got - exact) <= 1e-15 * max(abs(exact), 1.0)


def _worker_det_periodic(rank, world, port, q):
    """Reproducible builds of the periodic loops, sharded: PeriodicImplicitSim._allreduce_acc on [jh | j1] fp64 +
    [hi(2Ng) | lo(2Ng)] int64 and ExplicitSim.field_solve's all-reduce of the words behind rho_acc."""
    import types
    from pypic_b200 import fixedpoint as FP
    from pypic_b200.periodic import PeriodicImplicitSim
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = Comm()
        Ng, n = 16, 5000
        s = FP.scale_exponent_current1(-1.602e-19, 517.0, 25.85)
        rs = np.random.RandomState(12)
        node = rs.randint(0, 2 * Ng, n)
        v = rs.normal(0, 1e-9, n)
        mine = np.arange(n) % world == rank
        hi, lo = FP.split(v[mine], s)
        H = np.zeros(2 * Ng, dtype=np.int64); Lo = np.zeros(2 * Ng, dtype=np.int64)
        np.add.at(H, node[mine], hi); np.add.at(Lo, node[mine], lo)
        acc = torch.zeros(2 * Ng + 4 * Ng, dtype=torch.float64)
        acc[2 * Ng:].view(torch.int64).copy_(torch.as_tensor(np.concatenate([H, Lo])))
        PeriodicImplicitSim._allreduce_acc(types.SimpleNamespace(det=True, comm=comm, Ng=Ng, acc=acc))
        # the explicit loop's layout: fp64[g] + int64[2g] with g = Ng + 1 (the slice ExplicitSim.field_solve reduces)
        g = Ng + 1
        rho = torch.zeros(3 * g, dtype=torch.float64)
        rho[g:].view(torch.int64).copy_(torch.as_tensor(np.arange(2 * g, dtype=np.int64) * (rank + 1)))
        comm.allreduce_sum(rho[g:].view(torch.int64))
        q.put((rank, acc.numpy().copy(), rho.numpy().copy(), s))
        comm.barrier()
    finally:
        dist.destroy_process_group()


def test_world2_gloo_reproducible_periodic_builds_allreduce_the_words():
    import math
    from pypic_b200 import fixedpoint as FP
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_det_periodic, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Ng, n = 16, 5000
    s = res[0][3]
    assert res[0][1].tobytes() == res[1][1].tobytes() and res[0][2].tobytes() == res[1][2].tobytes()
    acc = res[0][1]
    assert np.array_equal(acc[:2 * Ng], np.zeros(2 * Ng))                  # the fp64 slots are filled by the field kernel's take
    words = acc[2 * Ng:].view(np.int64)
    rs = np.random.RandomState(12)
    node = rs.randint(0, 2 * Ng, n); v = rs.normal(0, 1e-9, n)
    for gi in range(2 * Ng):
        exact = math.fsum(v[node == gi].tolist())
        got = FP.merge(int(words[gi]), int(words[2 * Ng + gi]), s)
        assert abs(got - exact) <= 1e-15 * max(abs(exact), 1e-12)
    g = Ng + 1
    assert np.array_equal(res[0][2][g:].view(np.int64), np.arange(2 * g, dtype=np.int64) * 3)
