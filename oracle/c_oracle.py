"""TEST INFRASTRUCTURE ONLY -- builds and binds oracle/c/dd_oracle.c (plain-C
restatement of the sheath Picard timestep; see the header of that file).  Used by
tests/ as a scalable checker and by bench.py as the timed CPU baseline
(cpu_baseline.kind == "port").  Never imported by the product path."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "dd_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdd_oracle.so")
_lib = None


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", SRC, "-o", LIB, "-lm"]
        subprocess.run(cmd, check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        _lib.dd_picard_step_c.restype = C.c_int
        _lib.dd_picard_step_c.argtypes = [C.c_long, C.c_long, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                          dp, dp, dp, dp, dp, dp, C.c_double, C.c_int, dp, dp, dp, dp,
                                          C.POINTER(C.c_double), C.c_int]
        _lib.dd_oracle_max_threads.restype = C.c_int
        _lib.dd_oracle_team_size.restype = C.c_int
        _lib.dd_oracle_team_size.argtypes = [C.c_int]
    return _lib


def max_threads():
    return int(load().dd_oracle_max_threads())


def team_size(nthreads):
    """Threads OpenMP actually grants when `nthreads` are requested (what a timing should report)."""
    return int(load().dd_oracle_team_size(int(nthreads)))


def dd_picard_step(x0, u0, q2, m2, n_split, active, E0, p2c, Ng, dx, dt, L, tol, maxiter, nthreads=1):
    """active (fp64 1/0/-1) is mutated.  Returns x1,u1,E1,j1,k,r."""
    lib = load()
    N = len(x0)
    x1 = np.zeros(N); u1 = np.zeros(N); E1 = np.zeros(Ng); j1 = np.zeros(Ng)
    r = C.c_double()
    k = lib.dd_picard_step_c(N, int(n_split), int(Ng), dx, dt, L, p2c, np.ascontiguousarray(q2, dtype=np.float64),
                             np.ascontiguousarray(m2, dtype=np.float64), x0, u0, active, E0, tol, int(maxiter),
                             x1, u1, E1, j1, C.byref(r), int(nthreads))
    return x1, u1, E1, j1, k, r.value
