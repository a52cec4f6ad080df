import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200 import _lib, device as D
from pypic_b200.sheath import SheathSim
Ng = 4097; dx = 1e-5; dt = 1e-12; L = dx * (Ng - 1)
KB, ME, MP = 1.38E-23, 9.11E-31, 1.67E-27
kT = KB * 116000.
dev = torch.device("cuda", 0)
N = int(float(sys.argv[1])); sort = int(sys.argv[2]); nit = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sims = {}
for dep in ("warp", "window"):
    s = SheathSim(N, Ng, dx, dt, L * 1e19 / N, kBT=(kT, kT), carry_vw=False, deposit=dep, rng="philox", seed=1, device=dev, sort_every=8)
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    s.x0.uniform_(0.0, 1.0, generator=gen).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
    s.u0.normal_(0.0, 1.0, generator=gen)
    s.u0[:s.n_split].mul_(float(np.sqrt(kT / ME))); s.u0[s.n_split:].mul_(float(np.sqrt(kT / MP)))
    if sort:
        s.sort_by_cell()
    s.E0.normal_(0.0, 1e4, generator=gen)
    s.Es.copy_(s.E0)
    sims[dep] = s
a, b = sims["warp"], sims["window"]
assert torch.equal(a.x0, b.x0) and torch.equal(a.u0, b.u0) and torch.equal(a.Es, b.Es)
for it in range(nit):
    accs = {}
    for dep, s in sims.items():
        s.acc.zero_()
        _lib.call("pic_dev_dd_picard_iter", C.byref(s.params), D.ptr(s.x0), D.ptr(s.u0), D.ptr(s.x1), D.ptr(s.u1), D.ptr(s.active),
                  D.ptr(s.Es), D.ptr(s.acc), 1 if it == 0 else 0, D.ptr(s.range_err), D.stream())
        accs[dep] = s.acc.clone()
    torch.cuda.synchronize()
    dx1 = (a.x1 != b.x1); du1 = (a.u1 != b.u1); dact = (a.active != b.active)
    nbx, nbu, nba = int(dx1.sum()), int(du1.sum()), int(dact.sum())
    da = (accs["warp"] - accs["window"]).abs()
    scale = float(accs["warp"][:2 * Ng].abs().max())
    print("N %.1e sort %d it %d: x1 mismatches %d, u1 %d, active %d; acc rel diff %.2e at %d; counts %s vs %s; dead %d" %
          (N, sort, it, nbx, nbu, nba, float(da[:2 * Ng].max()) / scale, int(da[:2 * Ng].argmax()), accs["warp"][2 * Ng:].tolist(),
           accs["window"][2 * Ng:].tolist(), int((a.active != 1).sum())))
    if nbx:
        wall = torch.nonzero(dx1).flatten().cpu().numpy()
        rows = (wall % 1024) // 64; warps = (wall % 16384) // 1024; chunks = wall // 16384
        print("   mismatch histogram by row:", np.bincount(rows, minlength=16).tolist())
        print("   by chunk# in cta:", np.bincount(chunks // 148, minlength=9).tolist(), " distinct (chunk,warp,row):", len(set(zip(chunks.tolist(), warps.tolist(), rows.tolist()))))
        if it == 0:
            ax1 = a.x1.cpu().numpy(); au1 = a.u1.cpu().numpy(); bu1 = b.u1.cpu().numpy(); bx1_ = b.x1.cpu().numpy()
            for i in wall[::64][:10]:
                dxx = bx1_[i] - ax1[i]; duu = bu1[i] - au1[i]
                print("     i %d row %d lane %d: dx1 %.3e du1 %.3e ratio dx1/du1 %.3e (dt/2 = %.1e); x1-x0 %.3e" %
                      (i, (i % 1024) // 64, (i % 64) // 2, dxx, duu, dxx / duu if duu else float('nan'), dt / 2, ax1[i] - float(a.x0[i])))
            key = a.x0.cpu().numpy() + dt * a.u0.cpu().numpy()
            order = np.argsort(key); ks = key[order]
            for i in wall[::64][:12]:
                pos = np.searchsorted(ks, bx1_[i]); cand = [order[min(max(pp, 0), N - 1)] for pp in (pos - 1, pos)]
                j = min(cand, key=lambda c: abs(key[c] - bx1_[i]))
                ch_i, ch_j = i // 16384, j // 16384
                print("     i %d (chunk %d cta %d warp %d row %d lane %d) used the data of particle %d (chunk %d = chunk%+d, warp %d row %d lane %d) resid %.2e" %
                      (i, ch_i, ch_i % 148, (i % 16384) // 1024, (i % 1024) // 64, (i % 64) // 2, j, ch_j, ch_j - ch_i, (j % 16384) // 1024,
                       (j % 1024) // 64, (j % 64) // 2, abs(key[j] - bx1_[i])))
            lanes = (wall % 64) // 2
            print("   mismatch histogram by lane:", np.bincount(lanes, minlength=32).tolist())
            X0h = a.x0.cpu().numpy(); U0h = a.u0.cpu().numpy(); bx1 = b.x1.cpu().numpy()
            for i in wall[::64][:6]:
                obs = bx1[i]
                for off in (-1024, -512, -256, -192, -128, -64, 64, 128, 192, 256, 512, 1024):
                    j = i + off
                    if 0 <= j < N:
                        approx = X0h[j] + dt * U0h[j]
                        if abs(approx - obs) < 1e-7 * L:
                            print("     particle", i, "got the data of particle i%+d (x0+dt*u0=%.12g, observed %.12g)" % (off, approx, obs))
                if abs(X0h[i] + dt * U0h[i] - obs) < 1e-7 * L:
                    print("     particle", i, "own data but small difference")
        w = wall[:4]
        for i in w:
            ch, r = divmod(int(i), 16384); wp, r2 = divmod(r, 1024); row, r3 = divmod(r2, 64)
            print("   i", i, "chunk", ch, "cta", ch % 148, "chunk# in cta", ch // 148, "warp", wp, "row", row, "lane", r3 // 2,
                  "warp x1", float(a.x1[i]), "window x1", float(b.x1[i]), "x0", float(a.x0[i]), "act", int(a.active[i]), int(b.active[i]))
    if float(da[:2 * Ng].max()) / scale > 1e-10:
        bad = torch.nonzero(da[:2 * Ng] / scale > 1e-10).flatten().cpu().numpy()
        print("   bad acc entries:", bad[:24], len(bad))
    # keep both in lock-step: feed the reference state to the window sim for the next iteration
    b.x1.copy_(a.x1); b.u1.copy_(a.u1); b.active.copy_(a.active)
    # field: simple deterministic perturbation so that iterations differ
    a.Es.mul_(0.97); b.Es.copy_(a.Es)
