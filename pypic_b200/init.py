"""Device-side initialisers and the IEAD histogram (N1): thin host
wrappers over pic_dev_init_uniform_maxwellian / pic_dev_pypic_perturb_positions /
pic_dev_gc_iead_hist.

The reference initialises on the host with NumPy's legacy MT19937 (pypic.initialize_p
pypic.py:384-470, PIC_L_DD.initialize :223-314, Particle._initialize_6D pygcpic.py:277-304);
the drop-in modules keep those (draw-order parity).  At 1e8-1e9 particles host initialisation
plus the upload dominates start-up, so the device-resident drivers can fill their stores
directly: same distributions, Philox4x32-10 keyed by the GLOBAL particle index (identical
state for any sharding), statistical parity only.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, device as D

kb = 1.38E-23


def fill_uniform_maxwellian(x, vs, n_split, xlo, xhi, sigma, mean=(0.0, 0.0), seed=1, stream_id=0, global_offset=0):
    """x ~ U(xlo,xhi); vs = up to three velocity tensors (or None): the first ~ N(mean[s], sigma[s]),
    the others ~ N(0, sigma[s]) with s = species of the slot (index >= n_split)."""
    vs = list(vs) + [None] * (3 - len(vs))
    N = (x if x is not None else vs[0]).numel()
    _lib.call("pic_dev_init_uniform_maxwellian", D.ptr(x), D.ptr(vs[0]), D.ptr(vs[1]), D.ptr(vs[2]), N, int(n_split),
              float(xlo), float(xhi), C.byref((C.c_double * 2)(*sigma)), C.byref((C.c_double * 2)(*mean)), int(seed),
              int(stream_id), int(global_offset), D.stream())


def init_sheath(sim, seed=1):
    """PIC_L_DD.initialize('beam') distributions on the device: x ~ U(0,L), u,v,w ~ N(0, sqrt(kBT_s/m_s))."""
    sig = (float(np.sqrt(sim.kBT[0] / sim.m[0])), float(np.sqrt(sim.kBT[1] / sim.m[1])))
    n = sim.N
    vs = [sim.u0[:n]] + ([sim.v0[:n], sim.w0[:n]] if sim.carry_vw else [])
    fill_uniform_maxwellian(sim.x0[:n], vs, sim.n_split, 0.0, sim.L, sig, seed=seed, global_offset=sim.start)
    sim.active.fill_(1)
    sim.kernel_launches += 1


def init_pypic(sim, system, perturbation, Kp, Te, seed=1):
    """pypic.initialize_p on the device for the three shipped systems: velocities by system
    (pypic.py:425-455) and the cosine perturbation loader (:457-467).  Returns the growth rate
    expression's inputs unchanged to the caller (host scalars are computed by pypic.initialize_p)."""
    N, Ng, L, dx = sim.N_global, sim.Ng, sim.L, sim.dx
    n = sim.N
    kBTe = kb * Te
    me = 9.11E-31
    vt = float(np.sqrt(kBTe / me))
    x, v = sim.x0[:n], sim.v0[:n]
    if system == "landau-damping":
        # normal(0, v_thermal/sqrt(2)) with v_thermal = sqrt(2 kBTe/me)  ->  sigma = sqrt(kBTe/me)
        fill_uniform_maxwellian(x, [v], N, 0.0, L, (vt, vt), seed=seed, global_offset=sim.start)
    elif system == "two-stream":
        half = N // 2
        fill_uniform_maxwellian(x, [v], half, 0.0, L, (0.5 * vt, 0.5 * vt), mean=(-2.0 * vt, 2.0 * vt), seed=seed,
                                global_offset=sim.start)
        # n_split is GLOBAL for this call: shift it into the shard's local index space
        if sim.start:
            fill_uniform_maxwellian(x, [v], max(0, min(n, half - sim.start)), 0.0, L, (0.5 * vt, 0.5 * vt),
                                    mean=(-2.0 * vt, 2.0 * vt), seed=seed, global_offset=sim.start)
    elif system == "bump-on-tail":
        plasma = N * 5 // 6
        fill_uniform_maxwellian(x, [v], max(0, min(n, plasma - sim.start)), 0.0, L, (vt, vt / 20.), mean=(0.0, 4.0 * vt),
                                seed=seed, global_offset=sim.start)
    else:
        raise ValueError("unknown system %r" % (system,))
    X = np.linspace(0.0, L, Ng + 1)
    K = Kp * 2.0 * np.pi / L
    F = 1.0 + np.cos(K * X[:Ng])
    F = (N * perturbation) * F / np.sum(F)
    counts = F.astype(np.int64)                                   # int(F[i]) (pypic.py:462)
    prefix = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    d_prefix, d_X = D.to_dev(prefix, sim.dev, torch.int64), D.to_dev(X, sim.dev)   # keep alive across the launch
    _lib.call("pic_dev_pypic_perturb_positions", D.ptr(x), n, D.ptr(d_prefix), D.ptr(d_X), Ng, int(seed), int(sim.start),
              D.stream())
    torch.cuda.current_stream().synchronize()
    sim.kernel_launches += 2
    return int(prefix[-1])


def init_gc_store(store, grid, Ti, m, vx=0.0, seed=1):
    """Particle._initialize_6D (pygcpic.py:299-301): x ~ U(0, L), v ~ N(0, sqrt(kb*T/m)) + [vx,0,0]."""
    vth = float(np.sqrt(kb * Ti / m))
    n = store.N
    fill_uniform_maxwellian(store.r[0][:n], [store.r[3][:n], store.r[4][:n], store.r[5][:n]], n, 0.0, grid.length,
                            (vth, vth), mean=(float(vx), float(vx)), seed=seed)
    for c in (1, 2, 6):
        store.r[c].zero_()


def iead_histogram(store, select, Z_select, e_edges, a_edges, hist=None):
    """hist += np.histogram2d(kinetic_energy/e, angle_wrt_wall, (e_edges, a_edges)) of the particles with
    select==1 and Z==Z_select (pygcpic.py:1516-1527, 1574-1584).  Returns the device histogram
    (fp64 counts, shape (len(e_edges)-1, len(a_edges)-1))."""
    dev = store.dev
    ne, na = len(e_edges) - 1, len(a_edges) - 1
    if hist is None:
        hist = torch.zeros((ne, na), dtype=torch.float64, device=dev)
    de, da = D.to_dev(np.asarray(e_edges, dtype=np.float64), dev), D.to_dev(np.asarray(a_edges, dtype=np.float64), dev)
    _lib.call("pic_dev_gc_iead_hist", D.ptr(store.r[3]), D.ptr(store.r[4]), D.ptr(store.r[5]), D.ptr(store.m), D.ptr(select),
              D.ptr(store.Z) if Z_select is not None else None, int(Z_select or 0), store.N, D.ptr(de), ne, D.ptr(da), na,
              D.ptr(hist), D.stream())
    torch.cuda.current_stream().synchronize()
    return hist
