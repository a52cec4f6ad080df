"""Checkpoint / restart of the device-resident state.

The reference pickles its Python object lists every `checkpoint_saving` steps
(pygcpic.py:1378-1383, 1627-1632: `particles<t>.sav`, `grid<t>.sav`).  Here the state is a
structure of arrays in HBM, so a checkpoint is ONE `.npz` per rank holding those arrays as they
are (fp64 / int32 / int8, no pickling), the grid arrays, the scalar state and -- for runs that
must continue the reference's random stream -- the legacy MT19937 state of `numpy.random`.
Restoring uploads the arrays back; a restored run continues bit-identically (tested).

    save_sheath(sim, path) / load_sheath(sim, path)       pypic_b200.sheath.SheathSim
    save_gc(store, grid, path) / load_gc(path, device)    pypic_b200.gcstore.ParticleStore + GridDev
    to_reference_particles(store) -> list of dicts        what the reference's pickle would hold
"""
import numpy as np

FORMAT_VERSION = 1


def _rng_state():
    name, keys, pos, has_gauss, cached = np.random.get_state()
    return dict(rng_keys=np.asarray(keys, dtype=np.uint32), rng_pos=np.int64(pos), rng_has_gauss=np.int64(has_gauss),
                rng_cached=np.float64(cached))


def _set_rng_state(z):
    if "rng_keys" in z:
        np.random.set_state(("MT19937", z["rng_keys"], int(z["rng_pos"]), int(z["rng_has_gauss"]), float(z["rng_cached"])))


def _host(t, n=None):
    a = t.detach().cpu().numpy()
    return a if n is None else a[:n]


# ------------------------------------------------------------------ sheath (PIC_L_DD)
def sheath_state(sim):
    """Host copy of everything SheathSim needs to continue: particle arrays of this rank's
    shard, flags, fields, counters."""
    st = sim.download()                      # the reference's particle order, whatever the cell sort did
    d = dict(format=np.int64(FORMAT_VERSION), kind="sheath", N_global=np.int64(sim.N_global), start=np.int64(sim.start),
             stop=np.int64(sim.stop), n_split=np.int64(sim.n_split), Ng=np.int64(sim.Ng), dx=np.float64(sim.dx),
             dt=np.float64(sim.dt), p2c=np.float64(sim.p2c), t=np.int64(sim.t), carry_vw=np.int64(sim.carry_vw),
             x0=st["x0"], u0=st["u0"], active=st["active"].astype(np.int8), E0=st["E0"], j0=st["j0"])
    if sim.carry_vw:
        d["v0"] = st["v0"]; d["w0"] = st["w0"]
    d.update(_rng_state())
    return d


def save_sheath(sim, path):
    np.savez(path, **sheath_state(sim))


def load_sheath(sim, path, restore_rng=True):
    """Restores a checkpoint written by save_sheath into a SheathSim built with the same
    sizes and sharding."""
    import torch
    z = np.load(path, allow_pickle=False)
    if str(z["kind"]) != "sheath" or int(z["format"]) != FORMAT_VERSION:
        raise ValueError("not a sheath checkpoint of format %d: %s" % (FORMAT_VERSION, path))
    for k, want in (("N_global", sim.N_global), ("start", sim.start), ("stop", sim.stop), ("Ng", sim.Ng),
                    ("n_split", sim.n_split)):
        if int(z[k]) != int(want):
            raise ValueError("checkpoint %s=%d does not match the simulation (%d)" % (k, int(z[k]), int(want)))
    for k, want in (("dx", sim.dx), ("dt", sim.dt), ("p2c", sim.p2c)):
        if float(z[k]) != float(want):
            raise ValueError("checkpoint %s=%r does not match the simulation (%r)" % (k, float(z[k]), float(want)))
    if sim.carry_vw and "v0" not in z:
        raise ValueError("the simulation carries v,w but the checkpoint was written without them")
    n = sim.N
    # the shard's arrays are stored in the reference's particle order: SheathSim.upload slices global
    # arrays, so place them at the shard's offset of an (otherwise unread) global-length view
    class _Shard:
        def __init__(self, a, start):
            self.a, self.start = a, start

        def __getitem__(self, s):
            return self.a[s.start - self.start:s.stop - self.start]
    sh = lambda name: _Shard(z[name], sim.start)
    sim.upload(sh("x0"), sh("u0"), sh("v0") if sim.carry_vw else None, sh("w0") if sim.carry_vw else None,
               E0=z["E0"], active=sh("active"))
    sim.j0.copy_(torch.as_tensor(z["j0"]))
    sim.t = int(z["t"])
    if restore_rng:
        _set_rng_state(z)
    return sim


# ------------------------------------------------------------------ pygcpic store + grid
def gc_state(store, grid=None):
    n = store.N
    d = dict(format=np.int64(FORMAT_VERSION), kind="pygcpic", N=np.int64(n), mode=np.int64(store.mode),
             B=np.asarray(store.B, dtype=np.float64), Eyz=np.asarray(store.Eyz, dtype=np.float64),
             r=np.stack([_host(c, n) for c in store.r], 1), charge_state=_host(store.charge_state, n), m=_host(store.m, n),
             p2c=_host(store.p2c, n), Z=_host(store.Z, n), active=_host(store.active, n), at_wall=_host(store.at_wall, n),
             from_wall=_host(store.from_wall, n))
    if grid is not None:
        d.update(grid_ng=np.int64(grid.ng), grid_length=np.float64(grid.length), grid_Te=np.float64(grid.Te),
                 grid_bc=str(grid.bc), grid_rho=_host(grid.rho), grid_phi=_host(grid.phi), grid_E=_host(grid.E),
                 grid_n=_host(grid.n), grid_state=_host(grid.state), grid_added=np.float64(grid.added_particles))
    d.update(_rng_state())
    return d


def save_gc(store, grid, path):
    np.savez(path, **gc_state(store, grid))


def load_gc(path, device=None, restore_rng=True):
    """Returns (ParticleStore, GridDev or None) rebuilt on the device."""
    import torch
    from .gcstore import GridDev, ParticleStore
    z = np.load(path, allow_pickle=False)
    if str(z["kind"]) != "pygcpic" or int(z["format"]) != FORMAT_VERSION:
        raise ValueError("not a pygcpic checkpoint of format %d: %s" % (FORMAT_VERSION, path))
    store = ParticleStore.from_arrays(z["r"], z["charge_state"], z["m"], z["p2c"], Z=z["Z"], active=z["active"],
                                      at_wall=z["at_wall"], from_wall=z["from_wall"], B=tuple(z["B"]), Eyz=tuple(z["Eyz"]),
                                      device=device)
    store.mode = int(z["mode"])
    grid = None
    if "grid_ng" in z:
        grid = GridDev(int(z["grid_ng"]), float(z["grid_length"]), float(z["grid_Te"]), str(z["grid_bc"]), device=device)
        for name in ("rho", "phi", "E", "n", "state"):
            getattr(grid, name).copy_(torch.as_tensor(z["grid_" + name]))
        grid.added_particles = float(z["grid_added"])
    if restore_rng:
        _set_rng_state(z)
    return store, grid


def to_reference_particles(state):
    """The per-particle records the reference's pickle holds (pygcpic.py:77-112), from a
    gc_state()/np.load dictionary: a list of dicts with r, m, charge_state, p2c, Z and flags."""
    out = []
    for i in range(int(state["N"])):
        out.append(dict(r=np.array(state["r"][i]), m=float(state["m"][i]), charge_state=float(state["charge_state"][i]),
                        p2c=float(state["p2c"][i]), Z=int(state["Z"][i]), active=int(state["active"][i]),
                        at_wall=int(state["at_wall"][i]), from_wall=int(state["from_wall"][i])))
    return out
