#!/usr/bin/env python
"""pygcpic.run_sheath with the particle list sharded over the ranks (torchrun) against the same list on one
rank: identical per-step integer outcomes (list length, wall hits, deletions, re-activations, ionisations,
mid-domain exits) and RNG consumption, particles equal to round-off (the all-reduced deposit is summed in
another order).  Two cases: a species-uniform hydrogen store on the fused push+deposit kernel, and a mixed
store (H+, H0, B0..2+, wall-born flags) with Monte-Carlo ionisation.
usage: torchrun --nproc-per-node 2 tools/gc_sharded_check.py [N]"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygcpic as G
from pypic_b200.dist import Comm, shard_range

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
comm = Comm()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 60000
# the physical parameters of the reference-generated mini-driver golden (tests/golden/gc.npz): 1e19 m^-3, Te = 50 eV
ng, Lg, Te, Ti, dt = 121, 0.0006646924998553118, 580000., 116000., 8e-11
B = np.array([0.13951295, 1.9951281, 0.])
p2c = Lg * 1e19 / N


def build(mixed, seed):
    rs = np.random.RandomState(seed)
    r = np.zeros((N, 7))
    r[:, 0] = rs.uniform(0, Lg, N)
    cs = np.ones(N); m = np.full(N, G.mp); pc = np.full(N, p2c); Z = np.ones(N, dtype=np.int32)
    fw = np.zeros(N, dtype=np.int8)
    if mixed:
        kind = rs.choice(3, N, p=[0.7, 0.15, 0.15])
        cs[kind == 1] = 0.; pc[kind == 1] = p2c * 2e-4; fw[kind == 1] = rs.randint(0, 2, int((kind == 1).sum()))
        nb = int((kind == 2).sum())
        cs[kind == 2] = rs.randint(0, 3, nb); m[kind == 2] = 10.81 * G.mp; pc[kind == 2] = p2c * 2e-4; Z[kind == 2] = 5
        fw[kind == 2] = rs.randint(0, 2, nb)
    r[:, 3:6] = rs.normal(0, 1.0, (N, 3)) * np.sqrt(G.kb * Ti / m)[:, None]
    active = np.ones(N, dtype=np.int8); active[rs.choice(N, N // 50, replace=False)] = 0
    return r, cs, m, pc, Z, fw, active


def run(mixed, sharded, steps=8):
    r, cs, m, pc, Z, fw, active = build(mixed, 7)
    a, b = shard_range(N, rank, world) if sharded else (0, N)
    np.random.seed(123)                                    # the SAME stream on every rank
    host_grid = G.Grid(ng, Lg, Te)
    grid = G.GridDev(ng, Lg, Te, comm=comm if sharded else None)
    st = G.ParticleStore.from_arrays(r[a:b], cs[a:b], m[a:b], pc[a:b], Z=Z[a:b], from_wall=fw[a:b], active=active[a:b], B=B)
    st.FUSED_MIN = 0 if not mixed else st.FUSED_MIN
    src = G.source_distribution_6D(host_grid, Ti, G.mp)
    out = G.run_sheath(grid, st, dt, steps, int(0.69 * N) if mixed else int(0.985 * N), src, p2c, G.mp,
                       ionize_Te=Te if mixed else None)
    nxt = float(np.random.uniform())
    rr, fl = st.r_host(), st.flags_host()
    csf = st.charge_state[:st.N].cpu().numpy()
    if sharded and world > 1:
        got = [None] * world
        dist.all_gather_object(got, (rr, fl["active"], csf))
        rr = np.concatenate([g[0] for g in got]); act = np.concatenate([g[1] for g in got]); csf = np.concatenate([g[2] for g in got])
    else:
        act = fl["active"]
    return out, nxt, rr, act, csf


res = {}
for mixed in (False, True):
    o1, n1, r1, a1, c1 = run(mixed, False)
    o2, n2, r2, a2, c2 = run(mixed, True)
    keys = ["length", "hits", "deleted", "reactivated"] + (["ionised_h", "ionised_b", "midexit"] if mixed else [])
    same = {k: list(map(int, o1[k])) == list(map(int, o2[k])) for k in keys}
    ek = all(len(x) == len(y) and (len(x) == 0 or np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 1e-9 * np.max(np.abs(x)))
             for x, y in zip(o1["ekin"], o2["ekin"]))
    n0 = float(np.max(np.abs(np.asarray(o1["n0"]) - np.asarray(o2["n0"])) / np.abs(o1["n0"])))
    shape_ok = r1.shape == r2.shape
    r_rel = float(np.max(np.abs(r1 - r2)) / np.max(np.abs(r1))) if shape_ok else None
    ok = all(same.values()) and ek and n1 == n2 and shape_ok and r_rel < 1e-9 and np.array_equal(a1, a2) and np.array_equal(c1, c2) and n0 < 1e-9
    res["mixed_ionising" if mixed else "uniform_fused"] = dict(
        ok=bool(ok), same=same, ekin_equal=bool(ek), rng_next_equal=n1 == n2, r_rel=r_rel, n0_rel=n0,
        steps=len(o1["length"]), length=list(map(int, o1["length"])), hits=list(map(int, o1["hits"])),
        reactivated=list(map(int, o1["reactivated"])), deleted=list(map(int, o1["deleted"])),
        ionised_h=list(map(int, o1["ionised_h"])) if mixed else None)
res["world"] = world; res["N"] = N
res["ok"] = all(v["ok"] for k, v in res.items() if isinstance(v, dict))
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
