// pypic.py (periodic implicit CN/Picard, electrons) and PIC_L.py (periodic explicit
// leapfrog with a Poisson solve every step).  Same design as dd_kernels.cu: the particle
// phase is one fused pass (gather + push + wrap + deposit) over the SoA store with the
// field / current tiles in shared memory.
#include "common.cuh"
#include "ring.cuh"
#include "host_common.h"

namespace pic {

template <bool AGG>
__device__ __forceinline__ void deposit2(double* tile, int iL, int iR, double vL, double vR, bool valid) {
    if (AGG) {
        unsigned full = 0xffffffffu;
        int key = valid ? iL : -1;
        int k0 = __shfl_sync(full, key, 0);
        if (__all_sync(full, key == k0)) {
            if (k0 < 0) return;
            vL = warp_sum(vL);
            vR = warp_sum(vR);
            if ((threadIdx.x & 31) == 0) { atomicAdd(&tile[iL], vL); atomicAdd(&tile[iR], vR); }
            return;
        }
    }
    if (valid) { atomicAdd(&tile[iL], vL); atomicAdd(&tile[iR], vR); }
}

// ============================================================ pypic.py
// Where a kernel's global additions go: fp64 REDs (fix == nullptr) or, in the reproducible build, integer atomics
// on fixed-point words (the sum is then independent of the order of the additions; see acc_add in dd_kernels.cu)
struct GAcc {
    double* acc;
    long long* fix;
    int nfix;
    double fs1;
    int* ferr;
    __device__ __forceinline__ void add(int n, double v) const {
        if (fix) {
            const double t = v * fs1, h = rint(t);
            if (!(fabs(h) < 4398046511104.0)) { if (ferr) atomicAdd(ferr, 1); return; }      // 2^42
            const long long lo = __double2ll_rn((t - h) * 4294967296.0);
            atomicAdd((unsigned long long*)fix + n, (unsigned long long)(long long)h);
            atomicAdd((unsigned long long*)fix + nfix + n, (unsigned long long)lo);
        } else {
            atomicAdd(&acc[n], v);
        }
    }
};
// reproducible build: fixed-point words [hi(n) | lo(n)] -> fp64 accumulator (one rounding per node), words cleared
__global__ void fix_take_k(double* __restrict__ acc, long long* __restrict__ fix, int n, double fi1) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const long long hi = fix[i], lo = fix[n + i];
        fix[i] = 0; fix[n + i] = 0;
        acc[i] += ((double)hi + (double)lo * (1.0 / 4294967296.0)) * fi1;
    }
}
static inline int fix_take_grid(int n) { int g = (n + 1023) / 1024; return g < 1 ? 1 : (g > 148 ? 148 : g); }

struct PYK {
    long long N;
    int Ng, flags;
    double dx, idx, dt, L, p2c, q, qm;
    const int* done;      // enqueue-ahead Picard loop: non-null and *done != 0 -> the particle kernels return at once
    // per-particle charge and mass (pypic.py:248 takes q, m as arrays: q_m = q / m): served by the
    // grid-stride kernel, which then reads them instead of the scalars above
    const double* qa;
    const double* ma;
    // reproducible build (flags bit7): fixed-point words [hi(2Ng) | lo(2Ng)] behind the fp64 accumulators [jh | j1]
    long long* fix;
    double fs1, fi1;
    int* ferr;            // error counter for contributions beyond the fixed-point range
};
static PYK make_pyk(const pic_pypic_params* p) {
    PYK k;
    k.done = nullptr; k.qa = nullptr; k.ma = nullptr;
    k.N = p->N; k.Ng = p->Ng; k.flags = p->flags; k.dx = p->dx; k.idx = 1. / p->dx; k.dt = p->dt;
    k.L = p->L; k.p2c = p->p2c; k.q = p->q; k.qm = p->q / p->m;
    k.fix = nullptr; k.fs1 = 1.0; k.fi1 = 1.0; k.ferr = nullptr;
    if (p->flags & 128) {
        // one contribution is q*v*p2c*w/dx with |v| below the speed of light (see make_ddk in dd_kernels.cu)
        int e = 0;
        frexp(fabs(p->q) * p->p2c * k.idx * 2.99792458e8, &e);
        k.fs1 = ldexp(1.0, 31 - e); k.fi1 = ldexp(1.0, e - 31);
    }
    return k;
}

__device__ __forceinline__ void pypic_fix(Cell& c, int Ng, int& bad) {
    if (c.iL < 0 || c.iL >= Ng || c.iR < 0 || c.iR >= Ng) {
        ++bad;                                   // x == L after the wrap etc.: reference reads out of bounds
        c.iL = clampi(c.iL, 0, Ng - 1);
        c.iR = clampi(c.iR, 0, Ng - 1);
    }
}

__global__ void pypic_interpolate_k(const double* __restrict__ F, const double* __restrict__ x,
                                    double* __restrict__ out, long long N, int Ng, double dx,
                                    int* __restrict__ range_err) {
    const double idx = 1. / dx;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_pypic<false>(x[i], dx, idx, Ng);
        pypic_fix(c, Ng, bad);
        out[i] = F[c.iL] * c.wL + F[c.iR] * c.wR;       // pypic.py:57
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

template <bool CURRENT>
__global__ void pypic_weight_k(const double* __restrict__ x, const double* __restrict__ q,
                               const double* __restrict__ v, double* __restrict__ acc, long long N, int Ng,
                               double dx, double p2c, int* __restrict__ range_err) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    const double idx = 1. / dx;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_pypic<!CURRENT>(x[i], dx, idx, Ng);
        pypic_fix(c, Ng, bad);
        double pre = CURRENT ? q[i] * v[i] * p2c * idx : q[i] * p2c * idx;   // pypic.py:121 / 168
        atomicAdd(&sm[c.iL], pre * c.wL);
        atomicAdd(&sm[c.iR], pre * c.wR);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Ng; i += blockDim.x)
        if (sm[i] != 0.0) atomicAdd(&acc[i], sm[i]);
    if (bad && range_err) atomicAdd(range_err, bad);
}

// weight_density_p / weight_current_p with order-independent accumulation (reproducible build: the initial rho of
// implicit_pic): one pair of fixed-point additions per particle on words [hi(Ng) | lo(Ng)]
template <bool CURRENT>
__global__ void pypic_weight_fix_k(const double* __restrict__ x, const double* __restrict__ q, const double* __restrict__ v,
                                   long long* __restrict__ fix, long long N, int Ng, double dx, double p2c, double fs1,
                                   int* __restrict__ range_err) {
    const double idx = 1. / dx;
    const GAcc ga = {nullptr, fix, Ng, fs1, range_err};
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_pypic<!CURRENT>(x[i], dx, idx, Ng);
        pypic_fix(c, Ng, bad);
        const double pre = CURRENT ? q[i] * v[i] * p2c * idx : q[i] * p2c * idx;   // pypic.py:121 / 168
        ga.add(c.iL, pre * c.wL);
        ga.add(c.iR, pre * c.wR);
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

// One Picard iteration of particle_push_p, particle phase (pypic.py:261-279).
// x1 holds the UNWRAPPED n+1 position of the previous iteration.
template <bool FIRST, bool AGG>
__global__ void __launch_bounds__(256) pypic_picard_iter_k(PYK k, const double* __restrict__ x0,
                                                           const double* __restrict__ v0, const double* x1i, double* x1,
                                                           double* __restrict__ v1, const double* __restrict__ Fs,
                                                           double* __restrict__ acc, int* __restrict__ range_err) {
    extern __shared__ double sm[];
    if (k.done && *(const volatile int*)k.done) return;
    const int Ng = k.Ng;
    double *sF = sm, *jh = sm + Ng, *j1 = sm + 2 * Ng;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) { sF[i] = Fs[i]; jh[i] = 0.0; j1[i] = 0.0; }
    __syncthreads();
    int bad = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nIter = (k.N + stride - 1) / stride;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double dtdt = k.dt * k.dt;
    for (long long itn = 0; itn < nIter; ++itn, i += stride) {
        bool valid = i < k.N;
        Cell ch, cf;
        ch.iL = ch.iR = cf.iL = cf.iR = 0;
        double hL = 0., hR = 0., fL = 0., fR = 0.;
        if (valid) {
            double X0 = ld_stream(x0 + i), V0 = ld_stream(v0 + i);
            if (k.flags & 2) X0 = wrap_mod(X0, k.L);                 // store keeps the unwrapped x1 of the last step
            double xs = FIRST ? X0 : wrap_mod((X0 + ld_stream(x1i + i)) * 0.5, k.L);
            Cell c = cell_pypic<false>(xs, k.dx, k.idx, Ng);
            pypic_fix(c, Ng, bad);
            double Ei = sF[c.iL] * c.wL + sF[c.iR] * c.wR;
            const double qi = k.qa ? k.qa[i] : k.q;
            const double qmi = k.qa ? qi / k.ma[i] : k.qm;           // q_m = q/m, pypic.py:248
            double X1 = X0 + k.dt * V0 + dtdt * qmi * Ei * 0.5;    // pypic.py:264
            double V1 = V0 + k.dt * qmi * Ei;                        // :265
            double XH = (X0 + X1) * 0.5, VH = (V0 + V1) * 0.5;      // :268-269
            st_stream(x1 + i, X1);
            if (!(k.flags & 8)) st_stream(v1 + i, V1);               // bit3: light iteration (no v1 store, no j1 deposit)
            double xhw = wrap_mod(XH, k.L);                          // :272
            double x1w = wrap_mod(X1, k.L);                          // :277
            ch = cell_pypic<false>(xhw, k.dx, k.idx, Ng);
            pypic_fix(ch, Ng, bad);
            double jh_i = qi * VH * k.p2c * k.idx;                   // :121
            hL = jh_i * ch.wL; hR = jh_i * ch.wR;
            if (!(k.flags & 8)) {       // bit3: light iteration, j1 (only used after the loop) is not deposited
                cf = cell_pypic<false>(x1w, k.dx, k.idx, Ng);
                pypic_fix(cf, Ng, bad);
                double j1_i = qi * V1 * k.p2c * k.idx;
                fL = j1_i * cf.wL; fR = j1_i * cf.wR;
            }
        }
        deposit2<AGG>(jh, ch.iL, ch.iR, hL, hR, valid);
        if (!(k.flags & 8)) deposit2<AGG>(j1, cf.iL, cf.iR, fL, fR, valid);
    }
    __syncthreads();
    for (int n = threadIdx.x; n < 2 * Ng; n += blockDim.x) {
        double v = sm[Ng + n];
        if (v != 0.0) atomicAdd(&acc[n], v);
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

// pypic.py:283-292 field phase (one CTA).  Es, Fs updated in place; acc zeroed.
__global__ void __launch_bounds__(1024) pypic_field_update_k(PYK k, double* __restrict__ acc,
                                                             const double* __restrict__ E0, double* __restrict__ Es,
                                                             double* __restrict__ Fs, double* __restrict__ E1,
                                                             double* __restrict__ j1o, double* __restrict__ stats,
                                                             double* __restrict__ Fs_prev, double* __restrict__ rhist,
                                                             int* __restrict__ ctl, double tol, int maxiter) {
    __shared__ double scratch[33];
    const int Ng = k.Ng;
    if (ctl && *(volatile int*)ctl) return;       // the loop already ended (enqueue-ahead mode)
    double* jh = acc;
    double* j1 = acc + Ng;
    double sh = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) { sh += jh[i]; s1 += j1[i]; j1o[i] = j1[i]; }
    sh = block_reduce<0>(sh, scratch);
    s1 = block_reduce<0>(s1, scratch);
    const double meanh = sh / (double)Ng;
    const double coef = k.dt / PIC_EPS0;
    double rr = 0.0, ee = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        int ip = (i + 1 == Ng) ? 0 : i + 1, im = (i == 0) ? Ng - 1 : i - 1;
        double sm_j = ((jh[ip] + 2.0 * jh[i]) + jh[im]) * 0.25;       // smooth_field_p(jh)
        double e0 = E0[i];
        double e1 = e0 + coef * (meanh - sm_j);                        // :283
        double eh = (e1 + e0) * 0.5;                                   // :285
        double d = Es[i] - eh;
        rr += d * d;                                                   // :289 (squared, no sqrt)
        ee += PIC_EPS0 * e1 * e1 * k.dx / 2.;
        E1[i] = e1;
        if (Fs_prev) Fs_prev[i] = Fs[i];   // the smoothed field this iteration gathered with
        Fs[i] = eh;        // staged: Eh, smoothed below
    }
    rr = block_reduce<0>(rr, scratch);
    ee = block_reduce<0>(ee, scratch);
    __syncthreads();
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) Es[i] = Fs[i];
    __syncthreads();
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        int ip = (i + 1 == Ng) ? 0 : i + 1, im = (i == 0) ? Ng - 1 : i - 1;
        Fs[i] = ((Es[ip] + 2.0 * Es[i]) + Es[im]) * 0.25;              // field for the next gather, :261
    }
    for (int i = threadIdx.x; i < 2 * Ng; i += blockDim.x) acc[i] = 0.0;
    if (threadIdx.x == 0) {
        const double it = stats[3] + 1.0;
        stats[0] = rr;
        stats[1] = s1 / (double)Ng;
        stats[2] = ee;
        stats[3] = it;
        if (rhist && it <= (double)maxiter) rhist[(int)it - 1] = rr;
        if (ctl && (!(rr > tol) || it >= (double)maxiter)) *ctl = 1;     // `while r > tol and k < maxiter`, pypic.py:259
    }
}

// Repair of a Picard loop that ended on a light iteration: deposit j1 = weight_current_p(x1 % L, q, v1)
// (pypic.py:277-279) from the committed x1, v1 into acc[Ng..2Ng) with the exact lookup.
__global__ void pypic_j1_repair_k(PYK k, const double* __restrict__ x0, const double* __restrict__ v0,
                                  const double* __restrict__ x1_prev, const double* __restrict__ x1,
                                  const double* __restrict__ Fs_prev, double* __restrict__ v1, int first,
                                  double* __restrict__ acc, int* __restrict__ range_err) {
    int bad = 0;
    const int Ng = k.Ng;
    const GAcc ga = {acc, k.fix, 2 * Ng, k.fs1, range_err};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        // v1 of the last (light) iteration: v0 + dt*(q/m)*E(xs) with the field and xs that iteration used (:261-265)
        double X0 = x0[i];
        if (k.flags & 2) X0 = wrap_mod(X0, k.L);
        const double xs = first ? X0 : wrap_mod((X0 + x1_prev[i]) * 0.5, k.L);
        Cell c = cell_pypic<false>(xs, k.dx, k.idx, Ng);
        pypic_fix(c, Ng, bad);
        const double Ei = Fs_prev[c.iL] * c.wL + Fs_prev[c.iR] * c.wR;
        const double V1 = v0[i] + k.dt * k.qm * Ei;
        v1[i] = V1;
        const double x1w = wrap_mod(x1[i], k.L);
        Cell cf = cell_pypic<false>(x1w, k.dx, k.idx, Ng);
        pypic_fix(cf, Ng, bad);
        const double j1_i = k.q * V1 * k.p2c * k.idx;
        ga.add(Ng + cf.iL, j1_i * cf.wL); ga.add(Ng + cf.iR, j1_i * cf.wR);
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}
// j1 part of pypic_field_update_k alone: j1 copied out, stats[1] = mean(j1), accumulator zeroed
__global__ void __launch_bounds__(1024) pypic_j1_finish_k(int Ng, double* __restrict__ acc, double* __restrict__ j1o,
                                                          double* __restrict__ stats) {
    __shared__ double scratch[33];
    double* j1 = acc + Ng;
    double s1 = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) { s1 += j1[i]; j1o[i] = j1[i]; }
    s1 = block_reduce<0>(s1, scratch);
    __syncthreads();
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) j1[i] = 0.0;
    if (threadIdx.x == 0) stats[1] = s1 / (double)Ng;
}

__global__ void wrap_periodic_k(double* __restrict__ x, long long N, double L) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        x[i] = wrap_mod(x[i], L);
}

// ============================================================ PIC_L.py
struct LK {
    long long N, n_split;
    int Ng, flags;
    double dx, idx, dt, L, p2c;
    double q[2], qm[2];
    // reproducible build (flags bit7, as pic_dd_params'): the global additions go to 2 x 64-bit fixed-point words
    // [hi(Ng+1) | lo(Ng+1)] behind the fp64 accumulator; one hi unit = 1/fs1
    long long* fix;
    double fs1, fi1;
    int* ferr;
};
static LK make_lk(const pic_l_params* p) {
    LK k;
    k.N = p->N; k.n_split = p->n_split; k.Ng = p->Ng; k.flags = p->flags; k.dx = p->dx; k.idx = 1. / p->dx;
    k.dt = p->dt; k.L = p->L; k.p2c = p->p2c;
    for (int s = 0; s < 2; ++s) { k.q[s] = p->q[s]; k.qm[s] = p->q[s] / p->m[s]; }
    k.fix = nullptr; k.fs1 = 1.0; k.fi1 = 1.0; k.ferr = nullptr;
    if (p->flags & 128) {
        // one contribution is q*p2c*w/dx with w <= 1: |v| < 2^e; a window column sums at most 2^10 of them per flush
        const double qa = fabs(p->q[0]) > fabs(p->q[1]) ? fabs(p->q[0]) : fabs(p->q[1]);
        int e = 0;
        frexp(qa * p->p2c * k.idx, &e);
        k.fs1 = ldexp(1.0, 31 - e); k.fi1 = ldexp(1.0, e - 31);
    }
    return k;
}
__device__ __forceinline__ void l_fix(Cell& c, int nodes, int& bad) {
    if (c.iL < 0 || c.iL > nodes - 2) { ++bad; c.iL = clampi(c.iL, 0, nodes - 2); c.iR = c.iL + 1; }
}

__global__ void l_interpolate_k(const double* __restrict__ F, const double* __restrict__ x, double* __restrict__ out,
                                long long N, int Ng, double dx, int* __restrict__ range_err) {
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_lper(x[i], dx, Ng + 1);
        l_fix(c, Ng + 1, bad);
        out[i] = c.wL * F[c.iL] + c.wR * F[c.iR];       // PIC_L.py:45
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

template <bool CURRENT>
__global__ void l_weight_k(const double* __restrict__ x, const double* __restrict__ q, const double* __restrict__ v,
                           double* __restrict__ acc, long long N, int Ng, double dx, double p2c,
                           int* __restrict__ range_err, int tile) {
    extern __shared__ double sm[];
    const int nodes = Ng + 1;
    double* dst = acc;                         // grids too large for a shared-memory tile: global REDs
    if (tile) {
        for (int i = threadIdx.x; i < nodes; i += blockDim.x) sm[i] = 0.0;
        __syncthreads();
        dst = sm;
    }
    const double idx = 1. / dx;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_lper(x[i], dx, nodes);
        l_fix(c, nodes, bad);
        // PIC_L.py:73-74 / 112-113: q*v*p2c*w*idx  |  q*p2c*w*idx
        double pre = CURRENT ? q[i] * v[i] * p2c : q[i] * p2c;
        atomicAdd(&dst[c.iL], pre * c.wL * idx);
        atomicAdd(&dst[c.iR], pre * c.wR * idx);
    }
    if (tile) {
        __syncthreads();
        for (int i = threadIdx.x; i < nodes; i += blockDim.x)
            if (sm[i] != 0.0) atomicAdd(&acc[i], sm[i]);
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}
// folds: rho[-1]=rho[0]+rho[-1]; rho[0]=rho[-1]  |  j[0]=j[-1]+j[0]; j[-1]=j[0]
__global__ void l_fold_k(const double* __restrict__ acc, double* __restrict__ out, int nodes, int current) {
    for (int i = threadIdx.x; i < nodes; i += blockDim.x) {
        double a = acc[i];
        if (i == 0 || i == nodes - 1) a = current ? (acc[nodes - 1] + acc[0]) : (acc[0] + acc[nodes - 1]);
        out[i] = a;
    }
}

// PIC_L.weightCurrents :48-60 / weightDensities :83-98: BOUNDED CIC on Ng nodes, no wall terms, no fold
template <bool CURRENT>
__global__ void l_weight_bounded_k(const double* __restrict__ x, const double* __restrict__ q, const double* __restrict__ v,
                                   double* __restrict__ acc, long long N, int Ng, double dx, double p2c,
                                   int* __restrict__ range_err) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    const double idx = 1. / dx;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_dd(x[i], dx);
        if (c.iL < 0 || c.iL > Ng - 2) { ++bad; c.iL = clampi(c.iL, 0, Ng - 2); }
        if (CURRENT) {          // (1./dx) * q[i] * p2c * v[i] * w   (:57-58)
            const double pre = idx * q[i] * p2c * v[i];
            atomicAdd(&sm[c.iL], pre * c.wL); atomicAdd(&sm[c.iL + 1], pre * c.wR);
        } else {                // q[i] * p2c * w * idx               (:93-94)
            const double pre = q[i] * p2c;
            atomicAdd(&sm[c.iL], pre * c.wL * idx); atomicAdd(&sm[c.iL + 1], pre * c.wR * idx);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Ng; i += blockDim.x)
        if (sm[i] != 0.0) atomicAdd(&acc[i], sm[i]);
    if (bad && range_err) atomicAdd(range_err, bad);
}
// PIC_L.pushParticlesImplicit :261-270: gather Eh at xh (periodic), xout = x0 + dt*v + dt*dt*(q/m)*E*0.5,
// vout = v + dt*(q/m)*E, per-particle q and m
__global__ void l_push_implicit_k(const double* __restrict__ x0, const double* __restrict__ xh, const double* __restrict__ v,
                                  const double* __restrict__ q, const double* __restrict__ m, const double* __restrict__ Eh,
                                  double* __restrict__ xout, double* __restrict__ vout, long long N, int Ng, double dx,
                                  double dt, int* __restrict__ range_err) {
    const int nodes = Ng + 1;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_lper(xh[i], dx, nodes);
        l_fix(c, nodes, bad);
        const double Ei = c.wL * Eh[c.iL] + c.wR * Eh[c.iR];
        const double qm = q[i] / m[i];
        xout[i] = x0[i] + dt * v[i] + dt * dt * qm * Ei * 0.5;
        vout[i] = v[i] + dt * qm * Ei;
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}
// PIC_L.applyBoundaryConditions :272-282: flag = 0 where x > L or x <= 0, 1 elsewhere (pic_dev_compact_flags mode 1)
__global__ void l_outside_flags_k(const double* __restrict__ x, int8_t* __restrict__ flags, long long N, double L) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double X = x[i];
        flags[i] = (X > L || X <= 0.) ? 0 : 1;
    }
}

// fused explicit particle phase: gather, kick-drift-kick, wrap, deposit rho(x_new)
template <bool AGG>
__global__ void __launch_bounds__(256) l_push_deposit_k(LK k, double* __restrict__ x, double* __restrict__ v,
                                                        const double* __restrict__ E, double* __restrict__ rho_acc,
                                                        int* __restrict__ range_err) {
    extern __shared__ double sm[];
    const int nodes = k.Ng + 1;
    double *sE = sm, *sR = sm + nodes;
    for (int i = threadIdx.x; i < nodes; i += blockDim.x) { sE[i] = E[i]; sR[i] = 0.0; }
    __syncthreads();
    int bad = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nIter = (k.N + stride - 1) / stride;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double hdt = k.dt * 0.5;
    const double wrapL = k.L + k.dx;
    for (long long itn = 0; itn < nIter; ++itn, i += stride) {
        bool valid = i < k.N;
        Cell cn;
        cn.iL = cn.iR = 0;
        double rL = 0., rR = 0.;
        if (valid) {
            int sp = i >= k.n_split;
            double X = ld_stream(x + i), V = ld_stream(v + i);
            Cell c = cell_lper(X, k.dx, nodes);
            l_fix(c, nodes, bad);
            double Ei = c.wL * sE[c.iL] + c.wR * sE[c.iR];
            double qm = sp ? k.qm[1] : k.qm[0];
            double vhalf = V + qm * hdt * Ei;            // PIC_L.py:255
            double xout = X + vhalf * k.dt;              // :256
            double vout = vhalf + qm * hdt * Ei;         // :257
            if (k.flags & 2) {                            // function-level pushParticlesExplicit: no wrap, no deposit
                st_stream(x + i, xout);
                st_stream(v + i, vout);
                valid = false;
            } else {
                double xw = wrap_mod(xout, wrapL);
                st_stream(x + i, xw);
                st_stream(v + i, vout);
                cn = cell_lper(xw, k.dx, nodes);
                l_fix(cn, nodes, bad);
                double pre = (sp ? k.q[1] : k.q[0]) * k.p2c;
                rL = pre * cn.wL * k.idx; rR = pre * cn.wR * k.idx;
            }
        }
        deposit2<AGG>(sR, cn.iL, cn.iR, rL, rR, valid);
    }
    __syncthreads();
    for (int n = threadIdx.x; n < nodes; n += blockDim.x)
        if (sR[n] != 0.0) atomicAdd(&rho_acc[n], sR[n]);
    if (bad && range_err) atomicAdd(range_err, bad);
}


// =====================================================================================
// v2 streaming kernels: the design of dd_picard_iter_v6_k (dd_kernels.cu) applied to the two
// periodic codes.  One persistent 512-thread CTA per SM; every warp owns contiguous slices of
// 1024 particles whose rows (64 particles of each input array) are staged through shared
// memory by 1-D TMA bulk copies into a per-warp ring (mbarrier completion), a lane handles two
// consecutive particles per row, and deposits go to the lane's PRIVATE window of S_W grid
// nodes in shared memory (conflict-free [node][thread] layout, plain LDS/DADD/STS).  The
// window is flushed with one column sum + S_W global REDs per slice.  The fast path assumes
// what holds for all but ~1e-5 of the particles of a store sorted by cell: position strictly
// inside the domain before and after the push (no periodic wrap), cell lookups not within
// 2^-20 of a cell edge (then floor(x*idx) is the true floor quotient and the remainder by one
// fma is exact), cell inside the warp's window.  Everything else is settled per particle by
// the exact routine of the v1 kernels with global REDs.
#define S_T 512
#define S_W 7
// wide build: 15-node windows (see V6_W_WIDE in dd_kernels.cu), chosen by the host when shared memory allows
#define S_W_WIDE 15
#define S_ROWS 16
#define S_CHUNK (S_T * 2 * S_ROWS)

template <int W>
__device__ __forceinline__ void swin_add(double* myw, double* __restrict__ acc, int wb, int c, double vL, double vR) {
    const unsigned d = (unsigned)(c - wb);
    if (d <= (unsigned)(W - 2)) { double* p = myw + d * S_T; p[0] += vL; p[S_T] += vR; }
    else { atomicAdd(&acc[c], vL); atomicAdd(&acc[c + 1], vR); }
}
template <int W>
__device__ __forceinline__ void swin_add(double* myw, const GAcc& ga, int wb, int c, double vL, double vR, int off = 0) {
    const unsigned d = (unsigned)(c - wb);
    if (d <= (unsigned)(W - 2)) { double* p = myw + d * S_T; p[0] += vL; p[S_T] += vR; }
    else { ga.add(off + c, vL); ga.add(off + c + 1, vR); }
}

// column sums of the warp's 32 private windows -> global REDs, one tile of W <= 16 columns per pass (one lane
// per column and half of the warp's windows); the windows are cleared
template <int TILES, int W>
__device__ __forceinline__ void swin_flush(double* win, double* myw, int wbase, int lane, int wb, double* __restrict__ acc,
                                           int tile_stride, int nodes) {
    static_assert(W >= 5 && W <= 16, "one lane per (column, half-warp)");
    const int n = lane >> 1, half = lane & 1;
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
        double s = 0.0;
        if (n < W) {
            const double* col = win + (t * W + n) * S_T + wbase + half * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) s += col[(j + n) & 15];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (n < W && half == 0) {
            const int node = wb + n;
            if (node >= 0 && node < nodes && s != 0.0) atomicAdd(&acc[t * tile_stride + node], s);
        }
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < TILES * W; ++n2) myw[n2 * S_T] = 0.0;
    __syncwarp();
}
// the same with the additions through GAcc (the column sums are fp64 sums in a fixed order: lanes, then rows -- for
// a given particle order they are reproducible; the merge into the global accumulator is what depends on scheduling)
template <int TILES, int W>
__device__ __forceinline__ void swin_flush(double* win, double* myw, int wbase, int lane, int wb, const GAcc& ga,
                                           int tile_stride, int nodes) {
    static_assert(W >= 5 && W <= 16, "one lane per (column, half-warp)");
    const int n = lane >> 1, half = lane & 1;
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
        double s = 0.0;
        if (n < W) {
            const double* col = win + (t * W + n) * S_T + wbase + half * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) s += col[(j + n) & 15];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (n < W && half == 0) {
            const int node = wb + n;
            if (node >= 0 && node < nodes && s != 0.0) ga.add(t * tile_stride + node, s);
        }
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < TILES * W; ++n2) myw[n2 * S_T] = 0.0;
    __syncwarp();
}

// ---------------------------------------------------------------- PIC_L explicit step
struct LFastC { double dx, idx, dt, qmh, qpi; unsigned hi_lim; };
struct LFastO { double X, V, fL, fR; int cF; unsigned fr, ps; bool emiss; };

// BIG (large-grid build): sE is the warp's window of L_EW field nodes starting at node eb; a gather cell outside
// it sets o.emiss and the particle is redone by the exact routine with the field read from global memory.
#define L_EW 32
template <bool BIG>
__device__ __forceinline__ void l_fast(const LFastC& c, const double* __restrict__ sE, int nodes, int eb, double X, double V,
                                       LFastO& o) {
    const double ts = X * c.idx, fs = floor(ts);
    const unsigned f0 = (unsigned)__double2hiint(ts - fs) - PIC_HI_G;
    const double rs = fma(-fs, c.dx, X);
    int is = min(max((int)fs, 0), nodes - 2);
    o.emiss = false;
    if (BIG) {
        is -= eb;
        o.emiss = (unsigned)is > (unsigned)(L_EW - 2);
        is = min(max(is, 0), L_EW - 2);
    }
    const double wR = div_const(rs, c.dx, c.idx), wL = 1.0 - wR;
    const double Ei = wL * sE[is] + wR * sE[is + 1];                 // PIC_L.py:45
    const double vh = V + c.qmh * Ei;                                 // :255
    o.X = X + vh * c.dt;                                              // :256
    o.V = vh + c.qmh * Ei;                                            // :257
    o.ps = max((unsigned)__double2hiint(X) - 1u, (unsigned)__double2hiint(o.X) - 1u);
    const double tf = o.X * c.idx, ff = floor(tf);
    const unsigned f1 = (unsigned)__double2hiint(tf - ff) - PIC_HI_G;
    const double rf = fma(-ff, c.dx, o.X);
    o.cF = (int)ff;
    o.fr = max(f0, f1);
    o.fR = c.qpi * (rf * c.idx); o.fL = c.qpi - o.fR;                 // :112-113 up to re-association
}

// exact per-particle routine (the body of l_push_deposit_k); deposits with global REDs
template <bool DET = false>
__device__ __noinline__ int l_particle_exact(const LK& k, long long i, double X, double V, const double* sE,
                                             double* __restrict__ rho_acc, double* x, double* v) {
    const int nodes = k.Ng + 1;
    const GAcc ga = {rho_acc, DET ? k.fix : nullptr, nodes, k.fs1, DET ? k.ferr : nullptr};
    int bad = 0;
    const int sp = i >= k.n_split;
    Cell c = cell_lper(X, k.dx, nodes);
    l_fix(c, nodes, bad);
    const double Ei = c.wL * sE[c.iL] + c.wR * sE[c.iR];
    const double qm = sp ? k.qm[1] : k.qm[0];
    const double hdt = k.dt * 0.5;
    const double vhalf = V + qm * hdt * Ei;
    const double xout = X + vhalf * k.dt;
    const double vout = vhalf + qm * hdt * Ei;
    const double xw = wrap_mod(xout, k.L + k.dx);
    x[i] = xw; v[i] = vout;
    Cell cn = cell_lper(xw, k.dx, nodes);
    l_fix(cn, nodes, bad);
    const double pre = (sp ? k.q[1] : k.q[0]) * k.p2c;
    ga.add(cn.iL, pre * cn.wL * k.idx);
    ga.add(cn.iR, pre * cn.wR * k.idx);
    return bad;
}

template <int NST, bool BIG = false, int W = S_W, bool DET = false>
__global__ void __launch_bounds__(S_T, 1) l_push_deposit_v2_k(const __grid_constant__ LK k, int nchunks_fr, double* x,
                                                               double* v, const double* __restrict__ E,
                                                               double* __restrict__ rho_acc, int* __restrict__ range_err) {
    extern __shared__ __align__(128) double sm[];
    __shared__ int s_bad;
    const int nodes = k.Ng + 1;
    // reproducible build: integer merges into the fixed-point words behind rho_acc (constant-folded away otherwise)
    const GAcc ga = {rho_acc, DET ? k.fix : nullptr, nodes, k.fs1, DET ? k.ferr : nullptr};
    const int NP = BIG ? (S_T / 32) * L_EW : ((nodes + 15) & ~15);
    const int nchunks = nchunks_fr & 0x0fffffff;
    const int FRm = BIG ? (int)((unsigned)nchunks_fr >> 28) : (S_ROWS - 1);      // rows per deposit / field window - 1
    double* sE = sm;
    double* win = sm + NP;                                   // [W][S_T]
    double* ring = win + W * S_T;                          // [warp][stage][x|v][64]
    unsigned long long* bars = (unsigned long long*)(ring + (S_T / 32) * NST * 128);
    if (!BIG) for (int i = threadIdx.x; i < nodes; i += S_T) sE[i] = E[i];
    double* const wE = sm + (threadIdx.x >> 5) * L_EW;       // BIG: this warp's field window
    const double* const fE = BIG ? wE : sE;                  // what the fast path gathers from
    const double* const gE = BIG ? E : sE;                   // what the exact routine gathers from
    int eb = 0;
    double* myw = win + threadIdx.x;
#pragma unroll
    for (int n = 0; n < W; ++n) myw[n * S_T] = 0.0;
    if (threadIdx.x == 0) s_bad = 0;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wbase = threadIdx.x & ~31;
    const int NOWIN = -0x40000000;
    const double* wring = ring + warp * (NST * 128);
    const uint32_t ring_s = smem_u32(wring);
    const uint32_t bar_s = smem_u32(bars + warp * NST);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    LFastC fc;
    fc.dx = k.dx; fc.idx = k.idx; fc.dt = k.dt;
    fc.hi_lim = (unsigned)__double2hiint(k.L + k.dx) - 1u;
    const double hdt = k.dt * 0.5;
    const int my_chunks = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long woff = (long long)warp * (64 * S_ROWS);
    const long long chunk_step = (long long)gridDim.x * S_CHUNK;
    auto issue = [&](long long base, int st) {
        if (elect_one()) {
            const uint32_t dst = ring_s + st * 1024, bar = bar_s + 8 * st;
            mbar_expect_tx(bar, 1024u);
            bulk_g2s(dst, x + base, 512, bar);
            bulk_g2s(dst + 512, v + base, 512, bar);
        }
    };
    long long cbase = (long long)blockIdx.x * S_CHUNK + woff;
    if (my_chunks > 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) issue(cbase + 64 * s, s);
    }
    int stage = 0, bad = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (int c = 0; c < my_chunks; ++c, cbase += chunk_step) {
        const bool mixed = cbase < k.n_split && cbase + 64 * S_ROWS > k.n_split;
        const bool sp_slice = cbase >= k.n_split;
        fc.qmh = (sp_slice ? k.qm[1] : k.qm[0]) * hdt;
        fc.qpi = (sp_slice ? k.q[1] : k.q[0]) * k.p2c * k.idx;
        const bool more = c + 1 < my_chunks;
        int wb = NOWIN;
        long long ci = cbase + 2 * lane;
#pragma unroll 1
        for (int row = 0; row < S_ROWS; ++row, ci += 64) {
            mbar_wait(bar_s + 8 * stage, phase);
            const double* sb = wring + stage * 128 + 2 * lane;
            const double2 X = *(const double2*)sb, V = *(const double2*)(sb + 64);
            const int st_cur = stage;
            if (++stage == NST) { stage = 0; phase ^= 1u; }
            bool straddle = false;
            if (mixed) {
                const long long rstart = cbase + 64 * row;
                const bool sp = rstart >= k.n_split;
                straddle = !sp && rstart + 64 > k.n_split;
                fc.qmh = (sp ? k.qm[1] : k.qm[0]) * hdt;
                fc.qpi = (sp ? k.q[1] : k.q[0]) * k.p2c * k.idx;
            }
            if (BIG && (row & FRm) == 0) {
                if (row > 0 && wb != NOWIN) { __syncwarp(); swin_flush<1, W>(win, myw, wbase, lane, wb, ga, 0, nodes); }
                const int cb = (int)floor(__shfl_sync(full, X.x, 0) * k.idx);
                eb = min(max(cb - L_EW / 4, 0), nodes - L_EW);
                __syncwarp();
                wE[lane] = __ldg(E + eb + lane);
                __syncwarp();
            }
            LFastO a, b;
            l_fast<BIG>(fc, fE, nodes, eb, X.x, V.x, a);
            l_fast<BIG>(fc, fE, nodes, eb, X.y, V.y, b);
            const bool ra = (a.fr > PIC_HI_SPAN) | (a.ps >= fc.hi_lim) | straddle | a.emiss;
            const bool rb = (b.fr > PIC_HI_SPAN) | (b.ps >= fc.hi_lim) | straddle | b.emiss;
            if ((row & FRm) == 0) {
                int nok = __reduce_add_sync(full, (ra ? 0 : 1) + (rb ? 0 : 1));
                int sum = __reduce_add_sync(full, (ra ? 0 : a.cF) + (rb ? 0 : b.cF));
                wb = nok ? sum / nok - (W - 2) / 2 : NOWIN;
            }
            if (!(ra | rb)) {
                __stcs((double2*)(x + ci), make_double2(a.X, b.X));
                __stcs((double2*)(v + ci), make_double2(a.V, b.V));
                swin_add<W>(myw, ga, wb, a.cF, a.fL, a.fR);
                swin_add<W>(myw, ga, wb, b.cF, b.fL, b.fR);
            } else {
                if (ra) bad += l_particle_exact<DET>(k, ci, X.x, V.x, gE, rho_acc, x, v);
                else { x[ci] = a.X; v[ci] = a.V; swin_add<W>(myw, ga, wb, a.cF, a.fL, a.fR); }
                if (rb) bad += l_particle_exact<DET>(k, ci + 1, X.y, V.y, gE, rho_acc, x, v);
                else { x[ci + 1] = b.X; v[ci + 1] = b.V; swin_add<W>(myw, ga, wb, b.cF, b.fL, b.fR); }
            }
            // refill the drained stage only after every lane's LDS of it has executed (see v6)
            __syncwarp();
            if (row < S_ROWS - NST) issue(cbase + 64 * (row + NST), st_cur);
            else if (more) issue(cbase + chunk_step + 64 * (row + NST - S_ROWS), st_cur);
        }
        __syncwarp();
        if (wb != NOWIN) swin_flush<1, W>(win, myw, wbase, lane, wb, ga, 0, nodes);
    }
    // the N % S_CHUNK particles behind the last whole chunk: at most one per thread, exact routine, global REDs
    for (long long i = (long long)nchunks * S_CHUNK + (long long)blockIdx.x * S_T + threadIdx.x; i < k.N; i += (long long)gridDim.x * S_T)
        bad += l_particle_exact<DET>(k, i, x[i], v[i], gE, rho_acc, x, v);
    if (bad) atomicAdd(&s_bad, bad);
    __syncthreads();
    if (threadIdx.x == 0 && s_bad && range_err) atomicAdd(range_err, s_bad);
}

// ---------------------------------------------------------------- pypic Picard iteration
struct PFastC { double dx, idx, dt, c1, c2, qpi, L; unsigned hi_lim; };
struct PFastO { double X1, V1, hL, hR, fL, fR; int cH, cF; unsigned fr, ps; bool emiss; };

// BIG (large-grid build, as dd_picard_iter_v6_k's): sF is the warp's window of PY_EW field nodes starting at node
// eb instead of the whole smoothed field; a gather cell outside it sets o.emiss and the particle is redone by the
// exact routine with the field read from global memory.
#define PY_EW 32
template <bool FIRST, bool J1, bool BIG>
__device__ __forceinline__ void py_fast(const PFastC& c, const double* __restrict__ sF, int Ng, int eb, double X0, double V0,
                                        double pX1, PFastO& o) {
    const double xs = FIRST ? X0 : (X0 + pX1) * 0.5;                  // wrapped xh of the previous iteration
    const double ts = xs * c.idx, fs = floor(ts);
    const unsigned f0 = (unsigned)__double2hiint(ts - fs) - PIC_HI_G;
    const double rs = fma(-fs, c.dx, xs);
    int is = min(max((int)fs, 0), Ng - 2);
    o.emiss = false;
    if (BIG) {
        is -= eb;
        o.emiss = (unsigned)is > (unsigned)(PY_EW - 2);
        is = min(max(is, 0), PY_EW - 2);
    }
    const double wR = rs * c.idx, wL = 1.0 - wR;                      // pypic.py:52-53
    const double Ei = sF[is] * wL + sF[is + 1] * wR;                  // :57
    o.X1 = X0 + c.dt * V0 + c.c2 * Ei * 0.5;                          // :264
    o.V1 = V0 + c.c1 * Ei;                                            // :265
    const double XH = (X0 + o.X1) * 0.5, VH = (V0 + o.V1) * 0.5;      // :268-269
    const unsigned p0 = (unsigned)__double2hiint(X0) - 1u, p1 = (unsigned)__double2hiint(o.X1) - 1u;
    o.ps = FIRST ? max(p0, p1) : __vimax3_u32(p0, p1, (unsigned)__double2hiint(pX1) - 1u);
    const double th = XH * c.idx, fh = floor(th);
    const unsigned f1 = (unsigned)__double2hiint(th - fh) - PIC_HI_G;
    const double rh = fma(-fh, c.dx, XH);
    o.cH = (int)fh;
    const double ah = c.qpi * VH;                                     // :121 up to re-association
    o.hR = ah * (rh * c.idx); o.hL = ah - o.hR;
    if (J1) {
        const double tf = o.X1 * c.idx, ff = floor(tf);
        const unsigned f2 = (unsigned)__double2hiint(tf - ff) - PIC_HI_G;
        const double rf = fma(-ff, c.dx, o.X1);
        o.cF = (int)ff;
        o.fr = __vimax3_u32(f0, f1, f2);
        const double af = c.qpi * o.V1;
        o.fR = af * (rf * c.idx); o.fL = af - o.fR;
    } else {
        o.cF = o.cH; o.fr = max(f0, f1); o.fR = 0.0; o.fL = 0.0;
    }
}

// exact per-particle routine (the body of pypic_picard_iter_k); deposits with global REDs
template <bool FIRST, bool DET = false>
__device__ __noinline__ int py_particle_exact(const PYK& k, long long i, double X0, double V0, double pX1,
                                              const double* sF, double* __restrict__ acc, double* x1, double* v1) {
    const int Ng = k.Ng;
    const GAcc ga = {acc, DET ? k.fix : nullptr, 2 * Ng, k.fs1, DET ? k.ferr : nullptr};
    int bad = 0;
    if (k.flags & 2) X0 = wrap_mod(X0, k.L);
    const double dtdt = k.dt * k.dt;
    const double xs = FIRST ? X0 : wrap_mod((X0 + pX1) * 0.5, k.L);
    Cell c = cell_pypic<false>(xs, k.dx, k.idx, Ng);
    pypic_fix(c, Ng, bad);
    const double Ei = sF[c.iL] * c.wL + sF[c.iR] * c.wR;
    const double X1 = X0 + k.dt * V0 + dtdt * k.qm * Ei * 0.5;
    const double V1 = V0 + k.dt * k.qm * Ei;
    const double XH = (X0 + X1) * 0.5, VH = (V0 + V1) * 0.5;
    x1[i] = X1;
    if (!(k.flags & 8)) v1[i] = V1;
    const double xhw = wrap_mod(XH, k.L), x1w = wrap_mod(X1, k.L);
    Cell ch = cell_pypic<false>(xhw, k.dx, k.idx, Ng);
    pypic_fix(ch, Ng, bad);
    const double jh_i = k.q * VH * k.p2c * k.idx;
    ga.add(ch.iL, jh_i * ch.wL); ga.add(ch.iR, jh_i * ch.wR);
    if (!(k.flags & 8)) {
        Cell cf = cell_pypic<false>(x1w, k.dx, k.idx, Ng);
        pypic_fix(cf, Ng, bad);
        const double j1_i = k.q * V1 * k.p2c * k.idx;
        ga.add(Ng + cf.iL, j1_i * cf.wL); ga.add(Ng + cf.iR, j1_i * cf.wR);
    }
    return bad;
}

template <bool FIRST, int NST, bool J1, bool BIG = false, int W = S_W, bool DET = false>
__global__ void __launch_bounds__(S_T, 1) pypic_picard_iter_v2_k(const __grid_constant__ PYK k, int nchunks_fr,
                                                                  const double* __restrict__ x0,
                                                                  const double* __restrict__ v0, const double* x1i, double* x1,
                                                                  double* v1, const double* __restrict__ Fs,
                                                                  double* __restrict__ acc, int* __restrict__ range_err) {
    extern __shared__ __align__(128) double sm[];
    __shared__ int s_bad;
    if (k.done && *(const volatile int*)k.done) return;
    constexpr int NA = FIRST ? 2 : 3;
    const int Ng = k.Ng;
    const GAcc ga = {acc, DET ? k.fix : nullptr, 2 * Ng, k.fs1, DET ? k.ferr : nullptr};      // reproducible build: integer merges
    // smoothed field: the whole grid or, in the large-grid build, one PY_EW-node window per warp
    const int NP = BIG ? (S_T / 32) * PY_EW : ((Ng + 15) & ~15);
    const int nchunks = nchunks_fr & 0x0fffffff;
    const int FRm = BIG ? (int)((unsigned)nchunks_fr >> 28) : (S_ROWS - 1);      // rows per deposit / field window - 1
    double* sF = sm;
    double* win = sm + NP;                                   // [2*W][S_T]
    double* ring = win + 2 * W * S_T;                      // [warp][stage][x0|v0|x1][64]
    unsigned long long* bars = (unsigned long long*)(ring + (S_T / 32) * NST * 192);
    if (!BIG) for (int i = threadIdx.x; i < Ng; i += S_T) sF[i] = Fs[i];
    double* const wE = sm + (threadIdx.x >> 5) * PY_EW;      // BIG: this warp's field window
    const double* const fE = BIG ? wE : sF;                  // what the fast path gathers from
    const double* const gE = BIG ? Fs : sF;                  // what the exact routine gathers from
    int eb = 0;
    double* myw = win + threadIdx.x;
#pragma unroll
    for (int n = 0; n < 2 * W; ++n) myw[n * S_T] = 0.0;
    if (threadIdx.x == 0) s_bad = 0;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wbase = threadIdx.x & ~31;
    const int NOWIN = -0x40000000;
    const double* wring = ring + warp * (NST * 192);
    const uint32_t ring_s = smem_u32(wring);
    const uint32_t bar_s = smem_u32(bars + warp * NST);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    PFastC fc;
    fc.dx = k.dx; fc.idx = k.idx; fc.dt = k.dt; fc.L = k.L;
    fc.c1 = k.dt * k.qm; fc.c2 = k.dt * k.dt * k.qm;
    fc.qpi = k.q * k.p2c * k.idx;
    // strictly inside (0, (Ng-1)*dx): no wrap and the right node is iL+1 (the last cell wraps to node 0)
    fc.hi_lim = (unsigned)__double2hiint((double)(Ng - 1) * k.dx) - 1u;
    const int my_chunks = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long woff = (long long)warp * (64 * S_ROWS);
    const long long chunk_step = (long long)gridDim.x * S_CHUNK;
    auto issue = [&](long long base, int st) {
        if (elect_one()) {
            const uint32_t dst = ring_s + st * 1536, bar = bar_s + 8 * st;
            mbar_expect_tx(bar, NA * 512u);
            bulk_g2s(dst, x0 + base, 512, bar);
            bulk_g2s(dst + 512, v0 + base, 512, bar);
            if (!FIRST) bulk_g2s(dst + 1024, x1i + base, 512, bar);
        }
    };
    long long cbase = (long long)blockIdx.x * S_CHUNK + woff;
    if (my_chunks > 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) issue(cbase + 64 * s, s);
    }
    int stage = 0, bad = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (int c = 0; c < my_chunks; ++c, cbase += chunk_step) {
        const bool more = c + 1 < my_chunks;
        int wb = NOWIN;
        long long ci = cbase + 2 * lane;
#pragma unroll 1
        for (int row = 0; row < S_ROWS; ++row, ci += 64) {
            mbar_wait(bar_s + 8 * stage, phase);
            const double* sb = wring + stage * 192 + 2 * lane;
            const double2 X0 = *(const double2*)sb, V0 = *(const double2*)(sb + 64);
            double2 pX1 = make_double2(0., 0.);
            if (!FIRST) pX1 = *(const double2*)(sb + 128);
            const int st_cur = stage;
            if (++stage == NST) { stage = 0; phase ^= 1u; }
            if (BIG && (row & FRm) == 0) {
                // few particles per cell: the windows are flushed and re-centred every FRm+1 rows, and the
                // field window is loaded around the gather cell of the row's first particle
                if (row > 0 && wb != NOWIN) { __syncwarp(); swin_flush<J1 ? 2 : 1, W>(win, myw, wbase, lane, wb, ga, Ng, Ng); }
                const double xf = __shfl_sync(full, FIRST ? X0.x : (X0.x + pX1.x) * 0.5, 0);
                const int cb = (int)floor(xf * k.idx);
                eb = min(max(cb - PY_EW / 4, 0), Ng - PY_EW);
                __syncwarp();
                wE[lane] = __ldg(Fs + eb + lane);
                __syncwarp();
            }
            PFastO a, b;
            py_fast<FIRST, J1, BIG>(fc, fE, Ng, eb, X0.x, V0.x, pX1.x, a);
            py_fast<FIRST, J1, BIG>(fc, fE, Ng, eb, X0.y, V0.y, pX1.y, b);
            const bool ra = (a.fr > PIC_HI_SPAN) | (a.ps >= fc.hi_lim) | a.emiss;
            const bool rb = (b.fr > PIC_HI_SPAN) | (b.ps >= fc.hi_lim) | b.emiss;
            if ((row & FRm) == 0) {
                int nok = __reduce_add_sync(full, (ra ? 0 : 1) + (rb ? 0 : 1));
                int sum = __reduce_add_sync(full, (ra ? 0 : a.cH) + (rb ? 0 : b.cH));
                wb = nok ? sum / nok - (W - 2) / 2 : NOWIN;
            }
            if (!(ra | rb)) {
                __stcs((double2*)(x1 + ci), make_double2(a.X1, b.X1));
                if (J1) __stcs((double2*)(v1 + ci), make_double2(a.V1, b.V1));
                swin_add<W>(myw, ga, wb, a.cH, a.hL, a.hR);
                if (J1) swin_add<W>(myw + W * S_T, ga, wb, a.cF, a.fL, a.fR, Ng);
                swin_add<W>(myw, ga, wb, b.cH, b.hL, b.hR);
                if (J1) swin_add<W>(myw + W * S_T, ga, wb, b.cF, b.fL, b.fR, Ng);
            } else {
                if (ra) bad += py_particle_exact<FIRST, DET>(k, ci, X0.x, V0.x, pX1.x, gE, acc, x1, v1);
                else {
                    x1[ci] = a.X1; if (J1) v1[ci] = a.V1;
                    swin_add<W>(myw, ga, wb, a.cH, a.hL, a.hR);
                    if (J1) swin_add<W>(myw + W * S_T, ga, wb, a.cF, a.fL, a.fR, Ng);
                }
                if (rb) bad += py_particle_exact<FIRST, DET>(k, ci + 1, X0.y, V0.y, pX1.y, gE, acc, x1, v1);
                else {
                    x1[ci + 1] = b.X1; if (J1) v1[ci + 1] = b.V1;
                    swin_add<W>(myw, ga, wb, b.cH, b.hL, b.hR);
                    if (J1) swin_add<W>(myw + W * S_T, ga, wb, b.cF, b.fL, b.fR, Ng);
                }
            }
            __syncwarp();
            if (row < S_ROWS - NST) issue(cbase + 64 * (row + NST), st_cur);
            else if (more) issue(cbase + chunk_step + 64 * (row + NST - S_ROWS), st_cur);
        }
        __syncwarp();
        if (wb != NOWIN) swin_flush<J1 ? 2 : 1, W>(win, myw, wbase, lane, wb, ga, Ng, Ng);
    }
    // the N % S_CHUNK particles behind the last whole chunk: at most one per thread, exact routine, global REDs
    for (long long i = (long long)nchunks * S_CHUNK + (long long)blockIdx.x * S_T + threadIdx.x; i < k.N; i += (long long)gridDim.x * S_T)
        bad += py_particle_exact<FIRST, DET>(k, i, x0[i], v0[i], FIRST ? 0.0 : x1i[i], gE, acc, x1, v1);
    if (bad) atomicAdd(&s_bad, bad);
    __syncthreads();
    if (threadIdx.x == 0 && s_bad && range_err) atomicAdd(range_err, s_bad);
}

// field phase build: fold rho, rhs of the gauge-fixed periodic system over `nodes` unknowns
// initial deposit of the reproducible build: rho of the current positions, one pair of fixed-point additions per particle
__global__ void l_weight_fix_k(LK k, const double* __restrict__ x, int* __restrict__ range_err) {
    const int nodes = k.Ng + 1;
    const GAcc ga = {nullptr, k.fix, nodes, k.fs1, range_err};
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_lper(x[i], k.dx, nodes);
        l_fix(c, nodes, bad);
        const double pre = (i >= k.n_split ? k.q[1] : k.q[0]) * k.p2c;
        ga.add(c.iL, pre * c.wL * k.idx);
        ga.add(c.iR, pre * c.wR * k.idx);
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}
__global__ void l_field_build_k(double* __restrict__ rho_acc, double* __restrict__ rho, double* __restrict__ a,
                                double* __restrict__ b, double* __restrict__ c, double* __restrict__ d, int nodes,
                                double dx) {
    __shared__ double scratch[33];
    double s = 0.0;
    double fold = rho_acc[0] + rho_acc[nodes - 1];
    for (int i = threadIdx.x; i < nodes; i += blockDim.x) {
        double r = (i == 0 || i == nodes - 1) ? fold : rho_acc[i];
        rho[i] = r;
        s += r;
    }
    s = block_reduce<0>(s, scratch);
    const double dx2 = dx * dx;
    const double c0 = -(s / (double)nodes) / PIC_EPS0;
    for (int i = threadIdx.x; i < nodes - 1; i += blockDim.x) {
        a[i] = 1.0; b[i] = -2.0; c[i] = 1.0;
        d[i] = -dx2 * c0 - dx2 * (rho[i] / PIC_EPS0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nodes; i += blockDim.x) rho_acc[i] = 0.0;
}
// phi = [x, 0] - max ; E = -dphi/dx (PIC_L.py:765-766, 235-246) ; stats[0] = sum(eps0 E^2/2)
__global__ void l_field_finish_k(const double* __restrict__ x, double* __restrict__ phi, double* __restrict__ E,
                                 int nodes, double dx, double* __restrict__ stats) {
    __shared__ double scratch[33];
    double m = -INFINITY;
    for (int i = threadIdx.x; i < nodes; i += blockDim.x) {
        double v = (i == nodes - 1) ? 0.0 : x[i];
        phi[i] = v;
        m = fmax(m, v);
    }
    m = block_reduce<1>(m, scratch);
    for (int i = threadIdx.x; i < nodes; i += blockDim.x) phi[i] = phi[i] - m;
    __syncthreads();
    double ee = 0.0;
    for (int i = threadIdx.x; i < nodes; i += blockDim.x) {
        double e;
        if (i == 0) e = -(phi[1] - phi[nodes - 1]) / dx * 0.5;
        else if (i == nodes - 1) e = -(phi[0] - phi[nodes - 2]) / dx * 0.5;
        else e = -(phi[i + 1] - phi[i - 1]) / dx * 0.5;
        E[i] = e;
        ee += PIC_EPS0 * e * e / 2.;
    }
    ee = block_reduce<0>(ee, scratch);
    if (threadIdx.x == 0 && stats) stats[0] = ee;
}

}  // namespace pic

using namespace pic;

extern "C" {

int pic_dev_pypic_interpolate(const double* F, const double* x, double* out, int64_t N, int Ng, double dx,
                              int* range_err, void* stream) {
    PIC_REQUIRE(F && x && out && N >= 0 && Ng >= 2, "pypic_interpolate: bad argument");
    if (N == 0) return PIC_OK;
    pypic_interpolate_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(F, x, out, N, Ng, dx, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_pypic_weight(const double* x, const double* q, const double* v, double* out, int64_t N, int Ng, double dx,
                         double p2c, int* range_err, void* stream) {
    PIC_REQUIRE(x && q && out && N >= 0 && Ng >= 2, "pypic_weight: bad argument");
    if (N == 0) return PIC_OK;
    size_t smem = (size_t)Ng * sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
    if (v) {
        PIC_CHECK_CUDA(cudaFuncSetAttribute(pypic_weight_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pypic_weight_k<true><<<grid_for(N, 256, 4), 256, smem, st>>>(x, q, v, out, N, Ng, dx, p2c, range_err);
    } else {
        PIC_CHECK_CUDA(cudaFuncSetAttribute(pypic_weight_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pypic_weight_k<false><<<grid_for(N, 256, 4), 256, smem, st>>>(x, q, v, out, N, Ng, dx, p2c, range_err);
    }
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_pypic_weight_fixed(const double* x, const double* q, const double* v, double* out, int64_t N, int Ng, double dx,
                               double p2c, double qmax, int* range_err, void* stream) {
    PIC_REQUIRE(x && q && out && N >= 0 && Ng >= 2 && qmax > 0, "pypic_weight_fixed: bad argument");
    if (N == 0) return PIC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int e = 0;
    frexp(qmax * p2c / dx * (v ? 2.99792458e8 : 1.0), &e);
    const double fs1 = ldexp(1.0, 31 - e), fi1 = ldexp(1.0, e - 31);
    long long* fix = nullptr;
    PIC_CHECK_CUDA(cudaMallocAsync((void**)&fix, (size_t)2 * Ng * sizeof(long long), st));
    PIC_CHECK_CUDA(cudaMemsetAsync(fix, 0, (size_t)2 * Ng * sizeof(long long), st));
    if (v) pypic_weight_fix_k<true><<<grid_for(N, 256, 8), 256, 0, st>>>(x, q, v, fix, N, Ng, dx, p2c, fs1, range_err);
    else pypic_weight_fix_k<false><<<grid_for(N, 256, 8), 256, 0, st>>>(x, q, v, fix, N, Ng, dx, p2c, fs1, range_err);
    PIC_CHECK_LAUNCH();
    fix_take_k<<<fix_take_grid(Ng), 1024, 0, st>>>(out, fix, Ng, fi1);
    PIC_CHECK_LAUNCH();
    PIC_CHECK_CUDA(cudaFreeAsync(fix, st));
    return PIC_OK;
}

static int pypic_iter_v1(const PYK& k, int flags, const double* x0, const double* v0, const double* x1i, double* x1,
                         double* v1, const double* Fs, double* acc, int first, int* range_err, cudaStream_t st) {
    size_t smem = (size_t)3 * k.Ng * sizeof(double);
    PIC_REQUIRE(smem <= (size_t)max_optin_smem() - 1024, "pypic_picard_iter: Ng too large for the shared-memory tiles");
    bool agg = !(flags & 1);
#define PIC_PY_LAUNCH(F, A)                                                                                     \
    do {                                                                                                        \
        auto kern = pypic_picard_iter_k<F, A>;                                                                  \
        PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        int occ = 0;                                                                                            \
        PIC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));                   \
        kern<<<grid_for(k.N, 256, occ > 0 ? occ : 1), 256, smem, st>>>(k, x0, v0, x1i, x1, v1, Fs, acc, range_err); \
    } while (0)
    if (first) { if (agg) PIC_PY_LAUNCH(true, true); else PIC_PY_LAUNCH(true, false); }
    else { if (agg) PIC_PY_LAUNCH(false, true); else PIC_PY_LAUNCH(false, false); }
#undef PIC_PY_LAUNCH
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

#define PY_NST 4
#define PY_NST_WIDE 3
// 15-node windows: default in the explicit kernel (+8 % per step with a sort every 16 steps instead of 8); in the
// Picard kernel they cost a ring stage and measure no gain (profiles/r2_periodic_wide.txt), so they are opt-in there
static bool s_narrow() {
    static const bool narrow = [] { const char* e = getenv("PIC_S_NARROW"); return e && e[0] == '1'; }();
    return narrow;
}
static bool s_wide_picard() {
    static const bool wide = [] { const char* e = getenv("PIC_S_WIDE_PICARD"); return e && e[0] == '1'; }();
    return wide;
}
int pic_dev_pypic_picard_iter(const pic_pypic_params* p, const double* x0, const double* v0, double* x1, double* v1,
                              const double* Fs, double* acc, int first, int* range_err, void* stream) {
    return pic_dev_pypic_picard_iter2(p, x0, v0, x1, x1, v1, Fs, acc, first, range_err, stream);
}

int pic_dev_pypic_picard_iter2(const pic_pypic_params* p, const double* x0, const double* v0, const double* x1i, double* x1,
                               double* v1, const double* Fs, double* acc, int first, int* range_err, void* stream) {
    return pic_dev_pypic_picard_iter3(p, x0, v0, x1i, x1, v1, Fs, acc, first, range_err, nullptr, stream);
}

int pic_dev_pypic_picard_iter3(const pic_pypic_params* p, const double* x0, const double* v0, const double* x1i, double* x1,
                               double* v1, const double* Fs, double* acc, int first, int* range_err, const int32_t* done_flag,
                               void* stream) {
    PIC_REQUIRE(p && x0 && v0 && x1i && x1 && v1 && Fs && acc, "pypic_picard_iter: null pointer");
    PIC_REQUIRE(p->Ng >= 2 && p->dx > 0, "pypic_picard_iter: bad parameters");
    if (p->N == 0) return PIC_OK;
    PYK k = make_pyk(p);
    k.done = done_flag;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem2 = ((size_t)((k.Ng + 15) & ~15) + (size_t)2 * S_W * S_T + (size_t)(S_T / 32) * PY_NST * 192 +
                          (size_t)(S_T / 32) * PY_NST) * sizeof(double);
    const bool aligned16 = (((uintptr_t)x0 | (uintptr_t)v0 | (uintptr_t)x1i | (uintptr_t)x1 | (uintptr_t)v1) & 15) == 0;
    long long done = 0;
    // reproducible build (flags bit7): acc is fp64[2Ng] followed by the fixed-point words int64[4Ng]; window kernels only
    const bool det = (p->flags & 128) != 0;
    if (det) {
        PIC_REQUIRE(!(p->flags & (1 | 4)) && aligned16 && k.Ng >= 8,
                    "pypic_picard_iter: the reproducible build needs the window kernel (16-byte aligned arrays, Ng >= 8)");
        k.fix = (long long*)(acc + 2 * k.Ng); k.ferr = range_err;
    }
    // large-grid build of the same kernel (flags bit4 forces it, for tests): per-warp field windows, so the
    // shared-memory footprint does not depend on Ng
    const size_t smem2b = ((size_t)(S_T / 32) * PY_EW + (size_t)2 * S_W * S_T + (size_t)(S_T / 32) * PY_NST * 192 +
                           (size_t)(S_T / 32) * PY_NST) * sizeof(double);
    const bool big = ((p->flags & 16) || smem2 > (size_t)max_optin_smem() - 512) && k.Ng >= PY_EW;
    if (big && !(p->flags & (1 | 4)) && aligned16 && (k.N >= S_CHUNK || det)) {
        const long long nchunks = k.N / S_CHUNK;
        PIC_REQUIRE(nchunks < (1 << 28), "pypic_picard_iter: shard too large");
        const bool light = (p->flags & 8) != 0;
        auto kern = first ? (light ? pypic_picard_iter_v2_k<true, PY_NST, false, true> : pypic_picard_iter_v2_k<true, PY_NST, true, true>)
                          : (light ? pypic_picard_iter_v2_k<false, PY_NST, false, true> : pypic_picard_iter_v2_k<false, PY_NST, true, true>);
        if (det)
            kern = first ? (light ? pypic_picard_iter_v2_k<true, PY_NST, false, true, S_W, true> : pypic_picard_iter_v2_k<true, PY_NST, true, true, S_W, true>)
                         : (light ? pypic_picard_iter_v2_k<false, PY_NST, false, true, S_W, true> : pypic_picard_iter_v2_k<false, PY_NST, true, true, S_W, true>);
        PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2b));
        // rows (of 64 particles) per deposit / field window: about three cells' worth of particles
        const double ppc = (double)k.N / (double)k.Ng;
        int fr = 16;
        while (fr > 1 && 64.0 * fr > 3.0 * ppc) fr >>= 1;
        long long cap = device_sm_count();
        const long long gridb = nchunks < cap ? (nchunks > 0 ? nchunks : 1) : cap;
        kern<<<(int)gridb, S_T, smem2b, st>>>(k, (int)((unsigned)nchunks | ((unsigned)(fr - 1) << 28)), x0, v0, x1i, x1, v1, Fs,
                                              acc, range_err);
        PIC_CHECK_LAUNCH();
        return PIC_OK;
    }
    if (!(p->flags & (1 | 4)) && aligned16 && k.Ng >= 8 && smem2 <= (size_t)max_optin_smem() - 512) {
        // default: TMA-staged private-window kernel over whole chunks, v1 kernel on the tail
        const long long nchunks = k.N / S_CHUNK;
        if (det) {
            const bool light = (p->flags & 8) != 0;
            auto kern = first ? (light ? pypic_picard_iter_v2_k<true, PY_NST, false, false, S_W, true> : pypic_picard_iter_v2_k<true, PY_NST, true, false, S_W, true>)
                              : (light ? pypic_picard_iter_v2_k<false, PY_NST, false, false, S_W, true> : pypic_picard_iter_v2_k<false, PY_NST, true, false, S_W, true>);
            PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            long long cap = device_sm_count();
            const long long grid = nchunks < cap ? (nchunks > 0 ? nchunks : 1) : cap;
            kern<<<(int)grid, S_T, smem2, st>>>(k, (int)nchunks, x0, v0, x1i, x1, v1, Fs, acc, range_err);
            PIC_CHECK_LAUNCH();
            return PIC_OK;
        }
        if (nchunks > 0) {
            const bool light = (p->flags & 8) != 0;
            // 15-node deposit windows (3 ring stages): env PIC_S_WIDE_PICARD=1, for A/B runs
            const size_t smem2w = ((size_t)((k.Ng + 15) & ~15) + (size_t)2 * S_W_WIDE * S_T + (size_t)(S_T / 32) * PY_NST_WIDE * 192 +
                                   (size_t)(S_T / 32) * PY_NST_WIDE) * sizeof(double);
            const bool wide = s_wide_picard() && smem2w <= (size_t)max_optin_smem() - 512;
            auto kern = wide ? (first ? (light ? pypic_picard_iter_v2_k<true, PY_NST_WIDE, false, false, S_W_WIDE>
                                               : pypic_picard_iter_v2_k<true, PY_NST_WIDE, true, false, S_W_WIDE>)
                                      : (light ? pypic_picard_iter_v2_k<false, PY_NST_WIDE, false, false, S_W_WIDE>
                                               : pypic_picard_iter_v2_k<false, PY_NST_WIDE, true, false, S_W_WIDE>))
                             : (first ? (light ? pypic_picard_iter_v2_k<true, PY_NST, false> : pypic_picard_iter_v2_k<true, PY_NST, true>)
                                      : (light ? pypic_picard_iter_v2_k<false, PY_NST, false> : pypic_picard_iter_v2_k<false, PY_NST, true>));
            const size_t smem = wide ? smem2w : smem2;
            PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            long long cap = device_sm_count();
            kern<<<(int)(nchunks < cap ? nchunks : cap), S_T, smem, st>>>(k, (int)nchunks, x0, v0, x1i, x1, v1, Fs, acc, range_err);
            PIC_CHECK_LAUNCH();
            return PIC_OK;                       // the kernel finishes the ragged tail itself
        }
        done = nchunks * S_CHUNK;
        if (done >= k.N) return PIC_OK;
    }
    PIC_REQUIRE(!det, "pypic_picard_iter: grid too large for the reproducible build's field tile and too small for the large-grid kernel");
    PYK t = k;
    t.N = k.N - done;
    return pypic_iter_v1(t, p->flags, x0 + done, v0 + done, x1i + done, x1 + done, v1 + done, Fs, acc, first, range_err, st);
}

int pic_dev_pypic_picard_iter_qm(const pic_pypic_params* p, const double* x0, const double* v0, const double* x1i,
                                 double* x1, double* v1, const double* q, const double* m, const double* Fs, double* acc,
                                 int first, int* range_err, const int32_t* done_flag, void* stream) {
    PIC_REQUIRE(p && x0 && v0 && x1i && x1 && v1 && q && m && Fs && acc, "pypic_picard_iter_qm: null pointer");
    PIC_REQUIRE(!(p->flags & 128), "pypic_picard_iter_qm: no reproducible build of the per-particle q, m kernel");
    PIC_REQUIRE(p->Ng >= 2 && p->dx > 0, "pypic_picard_iter_qm: bad parameters");
    PIC_REQUIRE(!(p->flags & 8), "pypic_picard_iter_qm: light iterations are not offered with per-particle q, m");
    if (p->N == 0) return PIC_OK;
    PYK k = make_pyk(p);
    k.done = done_flag; k.qa = q; k.ma = m;
    return pypic_iter_v1(k, p->flags | 4, x0, v0, x1i, x1, v1, Fs, acc, first, range_err, (cudaStream_t)stream);
}

int pic_dev_pypic_field_update(const pic_pypic_params* p, double* acc, const double* E0, double* Es, double* Fs,
                               double* E1, double* j1, double* stats, void* stream) {
    return pic_dev_pypic_field_update2(p, acc, E0, Es, Fs, E1, j1, stats, nullptr, nullptr, nullptr, 0.0, 0, stream);
}

int pic_dev_pypic_field_update2(const pic_pypic_params* p, double* acc, const double* E0, double* Es, double* Fs,
                                double* E1, double* j1, double* stats, double* Fs_prev, double* rhist, int32_t* ctl,
                                double tol, int maxiter, void* stream) {
    PIC_REQUIRE(p && acc && E0 && Es && Fs && E1 && j1 && stats, "pypic_field_update: null pointer");
    PIC_REQUIRE(!(ctl || rhist) || maxiter >= 1, "pypic_field_update: maxiter must be >= 1 with ctl / rhist");
    PYK k = make_pyk(p);
    if (p->flags & 128) {          // reproducible build: the currents arrive as fixed-point words behind acc
        // (a no-op launch of an ended loop finds zero words: the last real field update took them)
        fix_take_k<<<fix_take_grid(2 * k.Ng), 1024, 0, (cudaStream_t)stream>>>(acc, (long long*)(acc + 2 * k.Ng), 2 * k.Ng, k.fi1);
        PIC_CHECK_LAUNCH();
    }
    pypic_field_update_k<<<1, 1024, 0, (cudaStream_t)stream>>>(k, acc, E0, Es, Fs, E1, j1, stats, Fs_prev, rhist, ctl, tol,
                                                               maxiter);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_pypic_j1_repair(const pic_pypic_params* p, const double* x0, const double* v0, const double* x1_prev,
                            const double* x1_last, const double* Fs_prev, double* v1, int first, double* acc,
                            int* range_err, void* stream) {
    PIC_REQUIRE(p && x0 && v0 && x1_prev && x1_last && Fs_prev && v1 && acc, "pypic_j1_repair: null pointer");
    if (p->N == 0) return PIC_OK;
    PYK k = make_pyk(p);
    if (p->flags & 128) k.fix = (long long*)(acc + 2 * k.Ng);
    pypic_j1_repair_k<<<grid_for(k.N, 256, 8), 256, 0, (cudaStream_t)stream>>>(k, x0, v0, x1_prev, x1_last, Fs_prev, v1, first,
                                                                             acc, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}
int pic_dev_pypic_j1_finish(const pic_pypic_params* p, double* acc, double* j1, double* stats, void* stream) {
    PIC_REQUIRE(p && acc && j1 && stats, "pypic_j1_finish: null pointer");
    if (p->flags & 128) {
        PYK k = make_pyk(p);
        fix_take_k<<<fix_take_grid(2 * k.Ng), 1024, 0, (cudaStream_t)stream>>>(acc, (long long*)(acc + 2 * k.Ng), 2 * k.Ng, k.fi1);
        PIC_CHECK_LAUNCH();
    }
    pypic_j1_finish_k<<<1, 1024, 0, (cudaStream_t)stream>>>(p->Ng, acc, j1, stats);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_wrap_periodic(double* x, int64_t N, double L, void* stream) {
    PIC_REQUIRE(x && N >= 0 && L > 0, "wrap_periodic: bad argument");
    if (N == 0) return PIC_OK;
    wrap_periodic_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, N, L);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_l_interpolate(const double* F, const double* x, double* out, int64_t N, int Ng, double dx, int* range_err,
                          void* stream) {
    PIC_REQUIRE(F && x && out && N >= 0 && Ng >= 2, "l_interpolate: bad argument");
    if (N == 0) return PIC_OK;
    l_interpolate_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(F, x, out, N, Ng, dx, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_l_weight(const double* x, const double* q, const double* v, double* out, int64_t N, int Ng, double dx,
                     double p2c, int* range_err, void* stream) {
    PIC_REQUIRE(x && q && out && N >= 0 && Ng >= 2, "l_weight: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int nodes = Ng + 1;
    double* acc = nullptr;
    PIC_CHECK_CUDA(cudaMallocAsync((void**)&acc, (size_t)nodes * sizeof(double), st));
    PIC_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)nodes * sizeof(double), st));
    if (N > 0) {
        size_t smem = (size_t)nodes * sizeof(double);
        const int tile = smem <= (size_t)max_optin_smem() - 1024;
        if (!tile) smem = 0;
        if (v) {
            PIC_CHECK_CUDA(cudaFuncSetAttribute(l_weight_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem ? smem : 1)));
            l_weight_k<true><<<grid_for(N, 256, 4), 256, smem, st>>>(x, q, v, acc, N, Ng, dx, p2c, range_err, tile);
        } else {
            PIC_CHECK_CUDA(cudaFuncSetAttribute(l_weight_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem ? smem : 1)));
            l_weight_k<false><<<grid_for(N, 256, 4), 256, smem, st>>>(x, q, v, acc, N, Ng, dx, p2c, range_err, tile);
        }
        PIC_CHECK_LAUNCH();
    }
    l_fold_k<<<1, 1024, 0, st>>>(acc, out, nodes, v != nullptr);
    PIC_CHECK_LAUNCH();
    PIC_CHECK_CUDA(cudaFreeAsync(acc, st));
    return PIC_OK;
}

#define L_NST 6
int pic_dev_l_push_deposit(const pic_l_params* p, double* x, double* v, const double* E, double* rho_acc,
                           int* range_err, void* stream) {
    PIC_REQUIRE(p && x && v && E && rho_acc, "l_push_deposit: null pointer");
    PIC_REQUIRE(p->Ng >= 2 && p->dx > 0, "l_push_deposit: bad parameters");
    if (p->N == 0) return PIC_OK;
    LK k = make_lk(p);
    cudaStream_t st = (cudaStream_t)stream;
    const int nodes = k.Ng + 1;
    const size_t smem2 = ((size_t)((nodes + 15) & ~15) + (size_t)S_W * S_T + (size_t)(S_T / 32) * L_NST * 128 +
                          (size_t)(S_T / 32) * L_NST) * sizeof(double);
    const bool aligned16 = (((uintptr_t)x | (uintptr_t)v) & 15) == 0;
    long long done = 0;
    // reproducible build (flags bit7): rho_acc is fp64[nodes] followed by the fixed-point words int64[2*nodes]; only the
    // window kernels are built for it (their exact routine serves ragged tails and stores shorter than a chunk)
    const bool det = (p->flags & 128) != 0;
    if (det) {
        PIC_REQUIRE(!(p->flags & (1 | 2 | 4)) && aligned16, "l_push_deposit: the reproducible build needs the window kernel (16-byte aligned x, v)");
        k.fix = (long long*)(rho_acc + nodes); k.ferr = range_err;
    }
    // large-grid build of the same kernel (flags bit4 forces it, for tests): per-warp field windows
    const size_t smem2b = ((size_t)(S_T / 32) * L_EW + (size_t)S_W * S_T + (size_t)(S_T / 32) * L_NST * 128 +
                           (size_t)(S_T / 32) * L_NST) * sizeof(double);
    const bool big = ((p->flags & 16) || smem2 > (size_t)max_optin_smem() - 512) && nodes >= L_EW;
    if (big && !(p->flags & (1 | 2 | 4)) && aligned16 && (k.N >= S_CHUNK || det)) {
        const long long nchunks = k.N / S_CHUNK;
        PIC_REQUIRE(nchunks < (1 << 28), "l_push_deposit: shard too large");
        auto kern = det ? l_push_deposit_v2_k<L_NST, true, S_W, true> : l_push_deposit_v2_k<L_NST, true>;
        PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2b));
        const double ppc = (double)k.N / 2.0 / (double)nodes;       // two species interleave over the same cells
        int fr = 16;
        while (fr > 1 && 64.0 * fr > 3.0 * ppc) fr >>= 1;
        long long cap = device_sm_count();
        const long long gridb = nchunks < cap ? (nchunks > 0 ? nchunks : 1) : cap;
        kern<<<(int)gridb, S_T, smem2b, st>>>(k, (int)((unsigned)nchunks | ((unsigned)(fr - 1) << 28)), x, v, E, rho_acc, range_err);
        PIC_CHECK_LAUNCH();
        return PIC_OK;
    }
    if (!(p->flags & (1 | 2 | 4)) && aligned16 && k.Ng >= 8 && smem2 <= (size_t)max_optin_smem() - 512) {
        const long long nchunks = k.N / S_CHUNK;
        if (nchunks > 0 || det) {
            const size_t smem2w = smem2 + (size_t)(S_W_WIDE - S_W) * S_T * sizeof(double);      // 15-node windows
            const bool wide = !s_narrow() && smem2w <= (size_t)max_optin_smem() - 512;
            auto kern = det ? (wide ? l_push_deposit_v2_k<L_NST, false, S_W_WIDE, true> : l_push_deposit_v2_k<L_NST, false, S_W, true>)
                            : (wide ? l_push_deposit_v2_k<L_NST, false, S_W_WIDE> : l_push_deposit_v2_k<L_NST>);
            const size_t smem = wide ? smem2w : smem2;
            PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            long long cap = device_sm_count();
            const long long grid = nchunks < cap ? (nchunks > 0 ? nchunks : 1) : cap;
            kern<<<(int)grid, S_T, smem, st>>>(k, (int)nchunks, x, v, E, rho_acc, range_err);
            PIC_CHECK_LAUNCH();
            return PIC_OK;                       // the kernel finishes the ragged tail itself
        }
        done = nchunks * S_CHUNK;
        if (done >= k.N) return PIC_OK;
    }
    PIC_REQUIRE(!det, "l_push_deposit: the reproducible build needs a grid of at least 8 cells");
    LK t = k;
    t.N = k.N - done;
    t.n_split = k.n_split - done < 0 ? 0 : (k.n_split - done > t.N ? t.N : k.n_split - done);
    size_t smem = (size_t)2 * (k.Ng + 1) * sizeof(double);
    PIC_REQUIRE(smem <= (size_t)max_optin_smem() - 1024, "l_push_deposit: Ng too large for the shared-memory tiles");
    bool agg = !(p->flags & 1);
    auto kern = agg ? l_push_deposit_k<true> : l_push_deposit_k<false>;
    PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PIC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
    kern<<<grid_for(t.N, 256, occ > 0 ? occ : 1), 256, smem, st>>>(t, x + done, v + done, E, rho_acc, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_l_deposit_fixed(const pic_l_params* p, const double* x, double* rho_acc, int* range_err, void* stream) {
    PIC_REQUIRE(p && x && rho_acc, "l_deposit_fixed: null pointer");
    PIC_REQUIRE((p->flags & 128) && p->Ng >= 2 && p->dx > 0, "l_deposit_fixed: needs the parameters of the reproducible build (flags bit7)");
    if (p->N == 0) return PIC_OK;
    LK k = make_lk(p);
    k.fix = (long long*)(rho_acc + k.Ng + 1);
    l_weight_fix_k<<<grid_for(k.N, 256, 8), 256, 0, (cudaStream_t)stream>>>(k, x, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_l_field_solve(const pic_l_params* p, double* rho_acc, double* rho, double* phi, double* E, double* work,
                          double* stats, void* stream) {
    PIC_REQUIRE(p && rho_acc && rho && phi && E && work, "l_field_solve: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int nodes = p->Ng + 1;
    double *a = work, *b = work + nodes, *c = work + 2 * (size_t)nodes, *d = work + 3 * (size_t)nodes,
           *x = work + 4 * (size_t)nodes;
    if (p->flags & 128) {          // reproducible build: the deposits arrive as fixed-point words behind rho_acc
        LK k = make_lk(p);
        fix_take_k<<<fix_take_grid(nodes), 1024, 0, st>>>(rho_acc, (long long*)(rho_acc + nodes), nodes, k.fi1);
        PIC_CHECK_LAUNCH();
    }
    l_field_build_k<<<1, 1024, 0, st>>>(rho_acc, rho, a, b, c, d, nodes, p->dx);
    PIC_CHECK_LAUNCH();
    int rc = pic_dev_tridiag_pcr(a, b, c, d, x, nodes - 1, work + 5 * (size_t)nodes, stream);
    if (rc) return rc;
    l_field_finish_k<<<1, 1024, 0, st>>>(x, phi, E, nodes, p->dx, stats);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_l_weight_bounded(const double* x, const double* q, const double* v, double* out, int64_t N, int Ng, double dx,
                             double p2c, int* range_err, void* stream) {
    PIC_REQUIRE(x && q && out && N >= 0 && Ng >= 2, "l_weight_bounded: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    PIC_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)Ng * sizeof(double), st));
    if (N == 0) return PIC_OK;
    const size_t smem = (size_t)Ng * sizeof(double);
    PIC_REQUIRE(smem <= (size_t)max_optin_smem() - 1024, "l_weight_bounded: grid too large for the shared-memory tile");
    if (v) {
        PIC_CHECK_CUDA(cudaFuncSetAttribute(l_weight_bounded_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        l_weight_bounded_k<true><<<grid_for(N, 256, 4), 256, smem, st>>>(x, q, v, out, N, Ng, dx, p2c, range_err);
    } else {
        PIC_CHECK_CUDA(cudaFuncSetAttribute(l_weight_bounded_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        l_weight_bounded_k<false><<<grid_for(N, 256, 4), 256, smem, st>>>(x, q, v, out, N, Ng, dx, p2c, range_err);
    }
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_l_push_implicit(const double* x0, const double* xh, const double* v, const double* q, const double* m,
                            const double* Eh, double* xout, double* vout, int64_t N, int Ng, double dx, double dt,
                            int* range_err, void* stream) {
    PIC_REQUIRE(x0 && xh && v && q && m && Eh && xout && vout && N >= 0 && Ng >= 2, "l_push_implicit: bad argument");
    if (N == 0) return PIC_OK;
    l_push_implicit_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(x0, xh, v, q, m, Eh, xout, vout, N, Ng, dx, dt, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_l_outside_flags(const double* x, int8_t* flags, int64_t N, double L, void* stream) {
    PIC_REQUIRE(x && flags && N >= 0, "l_outside_flags: bad argument");
    if (N == 0) return PIC_OK;
    l_outside_flags_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, flags, N, L);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

}  // extern "C"
