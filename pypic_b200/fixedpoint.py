"""Host mirror of the fixed-point accumulation of the reproducible build
(`SheathSim(deposit="window-det")`, flags bit7 of pic_dd_params; csrc/dd_kernels.cu: make_ddk,
acc_add, fix_take).

A deposit v is split exactly into hi = rint(v * 2^s) and the remainder, which is rounded to a
multiple of 2^-(s+32); both words are accumulated with INTEGER additions, so the sum of a set of
deposits does not depend on the order in which warps, CTAs or ranks add them.  The scale is a
function of the run parameters only (the same on every rank): one deposit is q*p2c*u*w/dx with
|u| below the speed of light, hence |v| < 2^e and |hi| < 2^31 per deposit, which leaves room for
2^31 deposits per node in the 64-bit hi word.  These functions are used by the tests to check the
scheme (exactness of the split, size of the quantisation, order independence); the device code
does not call them.
"""
import math

import numpy as np

C_LIGHT = 2.99792458e8
LO_BITS = 32


def scale_exponent(q, p2c, dx):
    """s such that one hi unit is 2^-s (make_ddk: k.fs1 = 2^s)."""
    qa = max(abs(float(q[0])), abs(float(q[1])))
    _, e = math.frexp(qa * float(p2c) * (1.0 / float(dx)) * C_LIGHT)
    return 31 - e


def scale_exponent_density(q, p2c, dx):
    """The explicit loop's scale (make_lk in csrc/periodic_kernels.cu, pic_l_params.flags bit7): one deposit is
    q*p2c*w/dx with w <= 1; a window column sums at most 2^10 of them per flush, which the 2^42 guard of the hi
    word leaves room for."""
    qa = max(abs(float(q[0])), abs(float(q[1])))
    _, e = math.frexp(qa * float(p2c) * (1.0 / float(dx)))
    return 31 - e


def scale_exponent_current1(q, p2c, dx):
    """The periodic Picard loop's scale (make_pyk, pic_pypic_params.flags bit7): one species, deposits
    q*v*p2c*w/dx with |v| below the speed of light."""
    _, e = math.frexp(abs(float(q)) * float(p2c) * (1.0 / float(dx)) * C_LIGHT)
    return 31 - e


def split(v, s):
    """(hi, lo) integer words of the deposits v (acc_add)."""
    v = np.asarray(v, dtype=np.float64)
    t = np.ldexp(v, s)                      # exact: power-of-two scaling
    h = np.rint(t)
    lo = np.rint((t - h) * float(1 << LO_BITS))        # t - h is exact
    return h.astype(np.int64), lo.astype(np.int64)


def merge(hi, lo, s):
    """fp64 value of the accumulated words (fix_take): one rounding."""
    return (float(hi) + float(lo) * (1.0 / float(1 << LO_BITS))) * math.ldexp(1.0, -s)
