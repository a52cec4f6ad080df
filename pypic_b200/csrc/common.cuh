// Shared device helpers for the pyPIC B200 kernels (sm_100a only).
//
// Arithmetic rules (parity with the NumPy/Python reference):
//  * the library is compiled with -fmad=false, so a*b+c is NEVER contracted: every
//    per-particle expression rounds exactly like NumPy's unfused float64 ops;
//    fma() is used only where an exact residual is wanted (cell remainder).
//  * x/dx is the IEEE division (nvcc's double division is correctly rounded).
//  * x % dx (Python/NumPy floored modulo) is exact for x>=0; it is reproduced
//    without the slow fmod as r = fma(-k,dx,x) with k fixed up to the true
//    floor quotient (the true remainder is always representable, so with the
//    right k the fma is exact).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PIC_EPS0 8.854E-12
#define PIC_E    1.602E-19

namespace pic {

// Python/NumPy floored modulo for doubles (npy_divmod semantics).
__device__ __forceinline__ double py_mod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((b < 0.0) != (r < 0.0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

// x % L for the periodic wrap; fast exact paths for the common quotients.
__device__ __forceinline__ double wrap_mod(double x, double L) {
    if (x >= 0.0) {
        if (x < L) return x;
        if (x < 2.0 * L) return x - L;      // exact (Sterbenz)
    } else if (x > -L) {
        return x + L;                        // fmod(x,L)==x, then one rounded add
    }
    return py_mod(x, L);
}

// exact x % dx for x >= 0 given a candidate floor quotient kc (|kc - true| <= 1)
__device__ __forceinline__ double rem_exact(double x, double dx, double kc) {
    double r = fma(-kc, dx, x);
    if (r < 0.0) { kc -= 1.0; r = fma(-kc, dx, x); }
    else if (r >= dx) { kc += 1.0; r = fma(-kc, dx, x); }
    return r;
}

struct Cell { int iL; int iR; double wL; double wR; };

// ---- index/weight flavours, one per reference file ----------------------------
// PIC_L_DD.py:33-36,44-46 and pygcpic.py:344-346,873-876:
//   index = floor(x/dx) ; wR = (x % dx)/dx ; right node = index+1
__device__ __forceinline__ Cell cell_dd(double x, double dx) {
    Cell c;
    double qd = x / dx;
    double fl = floor(qd);
    double r = (x >= 0.0) ? rem_exact(x, dx, fl) : py_mod(x, dx);
    c.wR = r / dx;
    c.wL = 1.0 - c.wR;
    c.iL = (int)fl;
    c.iR = c.iL + 1;
    return c;
}

// a/b for a constant divisor b with y = RN(1/b) precomputed: one multiply and two
// Markstein corrections (exact residual by fma, then fma update).  After the first
// correction q is faithful; the second makes it the correctly rounded quotient, i.e.
// bit-identical to the IEEE division a/b (checked on the device against `/` by
// pic_dev_selftest_div, tests/test_gpu_math.py).  ~5 DP ops instead of ~12.
__device__ __forceinline__ double div_const(double a, double b, double y) {
    double q = a * y;
    double e = fma(-q, b, a);
    q = fma(e, y, q);
    e = fma(-q, b, a);
    return fma(e, y, q);
}

// Same result as cell_dd (bit for bit) without the two IEEE divisions on the hot path:
// floor(RN(x/dx)) equals floor(x*idx) unless x*idx is within a guard band of an integer,
// in which case the exact division decides (rare, <1e-8 of the particles).
__device__ __forceinline__ Cell cell_dd_fast(double x, double dx, double idx) {
    if (!(x >= 0.0)) return cell_dd(x, dx);
    Cell c;
    double t = x * idx;
    double fl = floor(t);
    double fr = t - fl;
    double g = 1e-9 + t * 1e-15;
    if (fr < g || fr > 1.0 - g) fl = floor(x / dx);
    double r = rem_exact(x, dx, fl);
    c.wR = div_const(r, dx, idx);
    c.wL = 1.0 - c.wR;
    c.iL = (int)fl;
    c.iR = c.iL + 1;
    return c;
}

// Deposit flavour of the lookup: EXACT cell index (same rule as cell_dd_fast) but the weight
// is r*idx (one multiply).  Deposited sums are re-associated by the parallel reduction
// anyway, so a last-bit difference in a single contribution is below the stated tolerance;
// per-particle state (gather -> x1,u1) always uses the exact weights of cell_dd_fast.
__device__ __forceinline__ int cell_dd_deposit(double x, double dx, double idx, double& wR) {
    if (!(x >= 0.0)) { Cell c = cell_dd(x, dx); wR = c.wR; return c.iL; }
    double t = x * idx;
    double fl = floor(t);
    double fr = t - fl;
    double g = 1e-9 + t * 1e-15;
    if (fr < g || fr > 1.0 - g) fl = floor(x / dx);
    double r = rem_exact(x, dx, fl);
    wR = r * idx;
    return (int)fl;
}

// PIC_L.py:40-43,103-106: index = floor(x/dx) % (Ng+1) ; right node = index+1
__device__ __forceinline__ Cell cell_lper(double x, double dx, int nodes /*Ng+1*/) {
    Cell c = cell_dd(x, dx);
    int i = c.iL % nodes;
    if (i < 0) i += nodes;
    c.iL = i;
    c.iR = i + 1;
    return c;
}

// pypic.py:45-53,110-118,157-165:
//   iL = int(x*(1/dx)) ; iR = int((x*(1/dx)+1) % Ng) ; wR = (x%dx)*idx  or  (x%dx)/dx
template <bool WEIGHT_BY_DIVISION>
__device__ __forceinline__ Cell cell_pypic(double x, double dx, double idx, int Ng) {
    Cell c;
    double t = x * idx;
    c.iL = (int)t;                              // truncation, like int(float)
    double tr = t + 1.0;
    double ng = (double)Ng;
    double m = (tr >= 0.0 && tr < ng) ? tr : ((tr >= ng && tr < 2.0 * ng) ? tr - ng : py_mod(tr, ng));
    c.iR = (int)m;
    double r = (x >= 0.0) ? rem_exact(x, dx, floor(t)) : py_mod(x, dx);
    c.wR = WEIGHT_BY_DIVISION ? (r / dx) : (r * idx);
    c.wL = 1.0 - c.wR;
    return c;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- block reductions (deterministic tree, blockDim multiple of 32, <=1024) -----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// all threads get the result; scratch must hold 33 doubles
template <int OP>  // 0 sum, 1 max, 2 min
__device__ __forceinline__ double block_reduce(double v, double* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = OP == 0 ? warp_sum(v) : (OP == 1 ? warp_max(v) : warp_min(v));
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        double ident = OP == 0 ? 0.0 : (OP == 1 ? -INFINITY : INFINITY);
        double t = lane < nw ? scratch[lane] : ident;
        t = OP == 0 ? warp_sum(t) : (OP == 1 ? warp_max(t) : warp_min(t));
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// streaming loads/stores: particle arrays are touched once per pass, keep them
// out of L1 so the field tiles / L2-resident grid data are not evicted.
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }

// ---- Philox4x32-10 counter RNG (device re-injection for benchmark-sized runs) -----
__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// uniform in (0,1) with 53 random bits
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    uint64_t v = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

}  // namespace pic
