"""CPU checks of the fixed-point accumulation scheme of the reproducible build (host mirror in
pypic_b200/fixedpoint.py of csrc/dd_kernels.cu: make_ddk / acc_add / fix_take)."""
import math

import numpy as np

from pypic_b200 import fixedpoint as FP

e, me, mp, kb = 1.602E-19, 9.11E-31, 1.67E-27, 1.38E-23


def _deposits(rs, n, sp, p2c, dx, kT):
    """CIC current deposits q*p2c*u*w/dx of n thermal particles of species sp."""
    m = (me, mp)[sp]
    q = (-e, e)[sp]
    u = rs.normal(0, 1, n) * math.sqrt(kT / m)
    w = rs.uniform(0, 1, n)
    return q * u * p2c * w / dx


def test_split_is_exact_to_the_lo_quantum_and_sum_is_order_independent():
    rs = np.random.RandomState(3)
    dx, p2c, kT = 1e-5, 2.048e9, 10 * e            # BASELINE config 2
    s = FP.scale_exponent((-e, e), p2c, dx)
    for sp in (0, 1):
        v = _deposits(rs, 200000, sp, p2c, dx, kT)
        hi, lo = FP.split(v, s)
        assert np.abs(hi).max() < 2 ** 31 and np.abs(lo).max() <= 2 ** 31
        # every deposit is reproduced to half a lo unit = 2^-(s+33)
        back = np.ldexp(hi.astype(np.float64), -s) + np.ldexp(lo.astype(np.float64), -s - 32)
        assert np.abs(back - v).max() <= math.ldexp(1.0, -s - 33) * (1 + 1e-12)
        # integer sums: any order, any partition into "ranks", gives the same words
        perm = rs.permutation(v.size)
        H, Lo = int(hi.sum()), int(lo.sum())
        assert int(hi[perm].sum()) == H and int(lo[perm].sum()) == Lo
        parts = np.array_split(perm, 8)
        assert sum(int(hi[p].sum()) for p in parts) == H and sum(int(lo[p].sum()) for p in parts) == Lo
        # and the merged value is the correctly rounded sum to within a few ulp of the RESULT
        # (fp64 atomics in a random order are only good to ~sqrt(n) ulp of the largest partial sum)
        exact = math.fsum(v.tolist())
        got = FP.merge(H, Lo, s)
        scale = math.fsum(np.abs(v).tolist())
        assert abs(got - exact) <= 2.0 ** -52 * abs(exact) + v.size * math.ldexp(1.0, -s - 33)
        # worst case of the quantisation: n half-units, ~2e-15 of sum|v| here -- four orders below the
        # worst-case bound n*2^-53*sum|v| of a sequential fp64 sum of the same deposits
        assert v.size * math.ldexp(1.0, -s - 33) < 1e-14 * scale


def test_scale_leaves_headroom_for_two_billion_deposits_per_node():
    for dx, p2c in ((1e-5, 2.048e9), (1e-5, 1.25e11), (4e-8, 1e6), (2.5e-3, 5170.0)):
        s = FP.scale_exponent((-e, e), p2c, dx)
        vmax = e * p2c / dx * FP.C_LIGHT          # |u| = c, weight 1
        hi, lo = FP.split(np.array([vmax, -vmax]), s)
        assert 2 ** 30 <= abs(int(hi[0])) < 2 ** 31 and int(hi[1]) == -int(hi[0])
        assert abs(int(hi[0])) * 2 ** 31 < 2 ** 63


def test_scales_of_the_periodic_loops():
    """make_lk / make_pyk (csrc/periodic_kernels.cu): a single deposit stays below 2^31 hi units, the largest window
    column sum (2^10 deposits) below the 2^42 guard, and the split is exact to the lo quantum."""
    rs = np.random.RandomState(4)
    # PIC_L.main's literals (density 1e10, N 1e5, Ng 200, dx 0.02) and the bench workload
    for dx, p2c in ((0.02, 4e5), (1e-5, 2.048e9)):
        s = FP.scale_exponent_density((-e, e), p2c, dx)
        vmax = e * p2c / dx
        hi, _ = FP.split(np.array([vmax]), s)
        assert 2 ** 30 <= abs(int(hi[0])) < 2 ** 31 and abs(int(hi[0])) * 2 ** 10 < 2 ** 42
        v = -e * p2c / dx * rs.uniform(0, 1, 100000)
        h, lo = FP.split(v, s)
        back = np.ldexp(h.astype(np.float64), -s) + np.ldexp(lo.astype(np.float64), -s - 32)
        assert np.abs(back - v).max() <= math.ldexp(1.0, -s - 33) * (1 + 1e-12)
        perm = rs.permutation(v.size)
        assert int(h[perm].sum()) == int(h.sum()) and int(lo[perm].sum()) == int(lo.sum())
    # pypic.main's literals (density 1e5, N 1e6, Ng 200): p2c truncated to an integer like numba's int32 argument
    L = 22.0 * math.sqrt(kb * 100.0 * 11600. * 8.854E-12 / e / e / 1e5)
    dx = L / 200.
    p2c = float(int(L * 1e5 / 1e6))
    s = FP.scale_exponent_current1(-e, max(p2c, 1.0), dx)
    hi, _ = FP.split(np.array([e * max(p2c, 1.0) / dx * FP.C_LIGHT]), s)
    assert 2 ** 30 <= abs(int(hi[0])) < 2 ** 31
