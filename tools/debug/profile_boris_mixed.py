"""Times the mixed-species Boris step (H+, H0, B+, B2+ in one list) on a B200: the fused kernel
(gc_push_boris_mix_k, push + walls + n and rho deposits in one pass) against the three v1 kernels it
replaces (push, apply_BCs, weight).  usage: profile_boris_mixed.py [N] [steps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pypic_b200.gcstore import GridDev, ParticleStore
MP = 1.67e-27
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ng = 4097; Lg = 4e-2; Te = 60 * 11600.; dt = 2e-11          # thermal hydrogen crosses ~0.15 cell per step
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)


def make(lean):
    st = ParticleStore(N, B=(2 * np.cos(1.5), 2 * np.sin(1.5), 0.), device=dev)
    st.r[0][:N].uniform_(0., 1., generator=g).mul_(Lg)
    for c in (3, 4, 5):
        st.r[c][:N].normal_(0., 7e4, generator=g)
    sp = torch.randint(0, 4, (N,), generator=g, device=dev)
    st.charge_state[:N] = torch.tensor([1., 0., 1., 2.], device=dev, dtype=torch.float64)[sp]
    st.m[:N] = torch.tensor([MP, MP, 10.81 * MP, 10.81 * MP], device=dev, dtype=torch.float64)[sp]
    st.p2c[:N] = torch.tensor([3.1e9, 1e9, 2e8, 2e8], device=dev, dtype=torch.float64)[sp]
    st.active[:N] = 1
    st.carry_yzt = not lean
    return st


for name, fused, lean in (("v1 kernels (push, BC, weight)", False, False), ("fused mixed, full store", True, False),
                          ("fused mixed, lean store", True, True)):
    grid = GridDev(ng, Lg, Te); grid.E.normal_(0., 5e3, generator=g)
    st = make(lean)
    st.sort_by_cell(grid)
    st.FUSED_MIN = 0 if fused else 10 ** 12

    def step():
        st.push_6D(dt, grid, deposit=fused)
        if fused:
            grid.finish_fused_deposit(1.0, dt)
        else:
            st.apply_BCs_dirichlet(grid)
            grid.weight_particles_to_grid_boltzmann(st, dt)
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        if i % 8 == 0:
            st.sort_by_cell(grid)
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # the push (+ fused deposit) launch alone, right after a sort
    st.sort_by_cell(grid)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(4):
        k0.record(); st.push_6D(dt, grid, deposit=fused); k1.record(); torch.cuda.synchronize()
        kms.append(k0.elapsed_time(k1))
        if fused:
            grid.finish_fused_deposit(1.0, dt)
    st.check(); grid.check()
    bpp = (7 * 16 + 3 * 8 + 1) if not lean else (4 * 16 + 3 * 8 + 1)
    print("%-32s %8.3f ms/step  %.3e particle-steps/s   push launch %.3f ms (%d B/particle -> %.0f GB/s)"
          % (name, ms, N / ms * 1e3, min(kms), bpp, N * bpp / min(kms) / 1e6))
    del st, grid
    torch.cuda.empty_cache()
