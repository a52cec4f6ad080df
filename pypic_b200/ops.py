"""NumPy-in / NumPy-out wrappers of the device kernels (function-level drop-ins).

Every function uploads its arguments, launches the CUDA kernel through the C ABI
and downloads the result -- the same contract as the reference's module-level
functions, which allocate and return fresh arrays.  No CPU arithmetic on the data.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, device as D


def _up(a, dev):
    return D.to_dev(np.asarray(a, dtype=np.float64), dev)


def _err(dev):
    return torch.zeros(1, dtype=torch.int32, device=dev)


# ------------------------------------------------------------------ grid kernels
def smooth(F, variant):
    dev = D.require_cuda()
    dF = _up(F, dev); out = torch.empty_like(dF)
    _lib.call("pic_dev_smooth", D.ptr(dF), D.ptr(out), dF.numel(), variant, D.stream())
    return out.cpu().numpy()


def differentiate(F, dx, variant):
    dev = D.require_cuda()
    dF = _up(F, dev); out = torch.empty_like(dF)
    _lib.call("pic_dev_differentiate", D.ptr(dF), D.ptr(out), dF.numel(), float(dx), variant, D.stream())
    return out.cpu().numpy()


def integrate_field(F, dx, subtract_max=False):
    dev = D.require_cuda()
    dF = _up(F, dev); out = torch.empty_like(dF)
    _lib.call("pic_dev_integrate_field", D.ptr(dF), D.ptr(out), dF.numel(), float(dx), int(subtract_max), D.stream())
    return out.cpu().numpy()


def tridiag(a, b, c, d):
    dev = D.require_cuda()
    da, db, dc, dd = (_up(v, dev) for v in (a, b, c, d))
    n = dd.numel()
    x = torch.empty_like(dd)
    work = D.f64(8 * n, dev) if n > _lib.PIC_PCR_SMEM_MAX else None
    _lib.call("pic_dev_tridiag_pcr", D.ptr(da), D.ptr(db), D.ptr(dc), D.ptr(dd), D.ptr(x), n, D.ptr(work), D.stream())
    return x.cpu().numpy()


def poisson_periodic(rho, dx, subtract_max=False):
    dev = D.require_cuda()
    dr = _up(rho, dev); n = dr.numel()
    phi = torch.empty_like(dr); work = D.f64(13 * n, dev)
    _lib.call("pic_dev_poisson_periodic", D.ptr(dr), D.ptr(phi), n, float(dx), int(subtract_max), D.ptr(work), D.stream())
    return phi.cpu().numpy()


def poisson_dirichlet(rho, dx):
    dev = D.require_cuda()
    dr = _up(rho, dev); n = dr.numel()
    phi = torch.empty_like(dr); work = D.f64(13 * n, dev)
    _lib.call("pic_dev_poisson_dirichlet", D.ptr(dr), D.ptr(phi), n, float(dx), D.ptr(work), D.stream())
    return phi.cpu().numpy()


def newton_boltzmann(src, phi_start, dx, n0, Te, bc, tol, iter_max):
    dev = D.require_cuda()
    ds = _up(src, dev); n = ds.numel()
    phi = _up(phi_start, dev) if phi_start is not None else D.f64(n, dev, True)
    it = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("pic_dev_newton_boltzmann", D.ptr(ds), D.ptr(phi), n, float(dx), float(n0), float(Te), int(bc),
              float(tol), int(iter_max), D.ptr(it), D.stream())
    return phi.cpu().numpy(), int(it.item())


def newton_boltzmann_l(rho, phi0, dx, kBT, tol, maxiter, periodic):
    """PIC_L / PIC_L_DD solvePoisson (periodic False) and solvePoissonPeriodic (True)."""
    dev = D.require_cuda()
    dr = _up(rho, dev); n = dr.numel()
    phi = _up(np.array(phi0, dtype=np.float64), dev)
    work = D.f64(n, dev, True)
    it = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("pic_dev_newton_boltzmann_l", D.ptr(dr), D.ptr(phi), n, float(dx), float(kBT), float(tol), int(maxiter),
              int(bool(periodic)), D.ptr(work), D.ptr(it), D.stream())
    return phi.cpu().numpy()


# ------------------------------------------------------------------ gathers
def _interp(name, F, x, Ng, dx):
    dev = D.require_cuda()
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    dF = _up(F, dev); dx_ = _up(x, dev); out = torch.empty_like(dx_); err = _err(dev)
    _lib.call(name, D.ptr(dF), D.ptr(dx_), D.ptr(out), x.size, int(Ng), float(dx), D.ptr(err), D.stream())
    D.check_range(err, name)
    return out.cpu().numpy()


def pypic_interpolate(F, x, Ng, dx):
    return _interp("pic_dev_pypic_interpolate", F, x, Ng, dx)


def dd_interpolate(F, x, Ng, dx):
    return _interp("pic_dev_dd_interpolate", F, x, Ng, dx)


def l_interpolate(F, x, Ng, dx):
    return _interp("pic_dev_l_interpolate", F, x, Ng, dx)


def gc_interpolate(E, x, ng, dx):
    return _interp("pic_dev_gc_interpolate", E, x, ng, dx)


# ------------------------------------------------------------------ deposits
def pypic_weight(x, q, v, p2c, Ng, dx, reproducible=False):
    """v=None -> weight_density_p, else weight_current_p; p2c truncated like numba's int32.
    reproducible: order-independent fixed-point accumulation (the same bits on every run)."""
    dev = D.require_cuda()
    dx_ = _up(x, dev); dq = _up(q, dev); dv = _up(v, dev) if v is not None else None
    out = D.f64(Ng, dev, True); err = _err(dev)
    if reproducible:
        qmax = float(np.max(np.abs(np.asarray(q)))) if np.size(q) else 1.0
        _lib.call("pic_dev_pypic_weight_fixed", D.ptr(dx_), D.ptr(dq), D.ptr(dv), D.ptr(out), dx_.numel(), int(Ng), float(dx),
                  float(int(p2c)), qmax if qmax > 0 else 1.0, D.ptr(err), D.stream())
    else:
        _lib.call("pic_dev_pypic_weight", D.ptr(dx_), D.ptr(dq), D.ptr(dv), D.ptr(out), dx_.numel(), int(Ng), float(dx),
                  float(int(p2c)), D.ptr(err), D.stream())
    D.check_range(err, "pypic weight")
    return out.cpu().numpy()


def dd_weight(x, q, v, p2c, Ng, dx, dt, active):
    dev = D.require_cuda()
    dx_ = _up(x, dev); dq = _up(q, dev); dv = _up(v, dev) if v is not None else None
    da = _up(active, dev); out = D.f64(Ng, dev, True); err = _err(dev)
    _lib.call("pic_dev_dd_weight", D.ptr(dx_), D.ptr(dq), D.ptr(dv), D.ptr(da), D.ptr(out), dx_.numel(), int(Ng),
              float(dx), float(dt), float(p2c), D.ptr(err), D.stream())
    D.check_range(err, "dd weight")
    return out.cpu().numpy()


def l_weight(x, q, v, p2c, Ng, dx):
    dev = D.require_cuda()
    dx_ = _up(x, dev); dq = _up(q, dev); dv = _up(v, dev) if v is not None else None
    out = D.f64(Ng + 1, dev, True); err = _err(dev)
    _lib.call("pic_dev_l_weight", D.ptr(dx_), D.ptr(dq), D.ptr(dv), D.ptr(out), dx_.numel(), int(Ng), float(dx),
              float(p2c), D.ptr(err), D.stream())
    D.check_range(err, "PIC_L weight")
    return out.cpu().numpy()


def l_weight_bounded(x, q, v, p2c, Ng, dx):
    """PIC_L.weightCurrents (v given) / weightDensities (v None): bounded CIC on Ng nodes."""
    dev = D.require_cuda()
    dx_ = _up(x, dev); dq = _up(q, dev); dv = _up(v, dev) if v is not None else None
    out = D.f64(Ng, dev, True); err = _err(dev)
    _lib.call("pic_dev_l_weight_bounded", D.ptr(dx_), D.ptr(dq), D.ptr(dv), D.ptr(out), dx_.numel(), int(Ng), float(dx),
              float(p2c), D.ptr(err), D.stream())
    D.check_range(err, "PIC_L bounded weight")
    return out.cpu().numpy()


def l_push_implicit(x0, xh, v, q, m, Ng, dt, dx, Eh):
    dev = D.require_cuda()
    a = [_up(t, dev) for t in (x0, xh, v, q, m, Eh)]
    xout = torch.empty_like(a[0]); vout = torch.empty_like(a[0]); err = _err(dev)
    _lib.call("pic_dev_l_push_implicit", *[D.ptr(t) for t in a], D.ptr(xout), D.ptr(vout), a[0].numel(), int(Ng), float(dx),
              float(dt), D.ptr(err), D.stream())
    D.check_range(err, "PIC_L.pushParticlesImplicit")
    return xout.cpu().numpy(), vout.cpu().numpy()


def l_outside(x, L):
    """Indices (ascending) of the particles with x > L or x <= 0 (PIC_L.applyBoundaryConditions)."""
    dev = D.require_cuda()
    dx_ = _up(x, dev); N = dx_.numel()
    flags = torch.empty(max(N, 1), dtype=torch.int8, device=dev)
    _lib.call("pic_dev_l_outside_flags", D.ptr(dx_), D.ptr(flags), N, float(L), D.stream())
    idx = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    bc = torch.zeros(2 * (N // 2048 + 2), dtype=torch.int64, device=dev)
    _lib.call("pic_dev_compact_flags", D.ptr(flags), N, 1, D.ptr(idx), D.ptr(cnt), D.ptr(bc), D.stream())
    return idx[:int(cnt.item())].cpu().numpy()


def gc_weight(x, charge_state, p2c, active, ng, dx):
    dev = D.require_cuda()
    dx_ = _up(x, dev); dcs = _up(charge_state, dev); dp = _up(p2c, dev)
    da = D.to_dev(np.asarray(active).astype(np.int8), dev, torch.int8)
    rho = D.f64(ng, dev, True); n = D.f64(ng, dev, True); err = _err(dev)
    _lib.call("pic_dev_gc_weight", D.ptr(dx_), D.ptr(dcs), D.ptr(dp), D.ptr(da), D.ptr(rho), D.ptr(n), dx_.numel(),
              int(ng), float(dx), D.ptr(err), D.stream())
    D.check_range(err, "gc weight")
    return rho.cpu().numpy(), n.cpu().numpy()


def compact_flags(flags, mode):
    dev = D.require_cuda()
    f = D.to_dev(np.asarray(flags).astype(np.int8), dev, torch.int8)
    N = f.numel()
    idx = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    bc = torch.zeros(2 * (N // 2048 + 2), dtype=torch.int64, device=dev)
    _lib.call("pic_dev_compact_flags", D.ptr(f), N, int(mode), D.ptr(idx), D.ptr(cnt), D.ptr(bc), D.stream())
    k = int(cnt.item())
    return idx[:k].cpu().numpy()
