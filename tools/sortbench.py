#!/usr/bin/env python
"""Times pic_dev_dd_sort_by_cell on a nearly sorted and on a random store."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypic_b200 import _lib, device as D
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000000
Ng = int(sys.argv[2]) if len(sys.argv) > 2 else 4097; dx = 1e-5; L = dx * (Ng - 1)
dev = torch.device("cuda", 0)
P = _lib.DDParams(N, N // 2, Ng, int(sys.argv[3]) if len(sys.argv) > 3 else 0, dx, 1e-12, L, 1.0, (C.c_double * 2)(0, 0), (C.c_double * 2)(1, 1))
x = torch.empty(N, dtype=torch.float64, device=dev).uniform_(0, L); u = torch.randn(N, dtype=torch.float64, device=dev)
xs = torch.empty_like(x); us = torch.empty_like(u)
cnt = torch.zeros(D.sort_counts_size(Ng), dtype=torch.int32, device=dev)
def run(a, b, c, d):
    _lib.call("pic_dev_dd_sort_by_cell", C.byref(P), D.ptr(a), D.ptr(b), None, None, D.ptr(c), D.ptr(d), None, None, D.ptr(cnt), D.stream())
def timed(a, b, c, d, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(a, b, c, d); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return ts
out = {"random": timed(x, u, xs, us)}
# check: sorted keys non-decreasing inside each species, multiset preserved
h = N // 2
k0 = torch.floor(xs[:h] / dx); k1 = torch.floor(xs[h:] / dx)
out["sorted_ok"] = bool((k0[1:] >= k0[:-1]).all().item() and (k1[1:] >= k1[:-1]).all().item())
out["sum_ok"] = bool(abs(float(xs.sum() - x.sum())) < 1e-6 * float(x.sum()))
# nearly sorted: perturb the sorted store by ~0.5 cell
x.copy_(xs); u.copy_(us)
x.add_(torch.randn(N, dtype=torch.float64, device=dev) * (0.5 * dx)).clamp_(1e-12, L * (1 - 1e-12))
out["nearly_sorted"] = timed(x, u, xs, us)
k0 = torch.floor(xs[:h] / dx); k1 = torch.floor(xs[h:] / dx)
out["sorted_ok2"] = bool((k0[1:] >= k0[:-1]).all().item() and (k1[1:] >= k1[:-1]).all().item())
# the stable radix sort of the reproducible build on the same nearly sorted store
scr = torch.zeros(D.sort_stable_scratch_size(N), dtype=torch.int32, device=dev)
where = C.c_int(0)
xa, ua = x.clone(), u.clone()
ts = []
for _ in range(3):
    xa.copy_(x); ua.copy_(u)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("pic_dev_dd_sort_by_cell_stable", C.byref(P), D.ptr(xa), D.ptr(ua), D.ptr(xs), D.ptr(us), D.ptr(scr),
              scr.numel(), C.byref(where), D.stream())
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
out["stable_nearly_sorted"] = ts
print(json.dumps(out))
