// Grid (field) kernels: binomial smoothing, finite differences, cumulative trapezoid,
// parallel-cyclic-reduction tridiagonal solves, Poisson and Newton-Boltzmann solves.
// All of them are O(Ng) work on <= a few MB: one CTA, shared memory, no tensor cores
// (a tridiagonal solve is not a dense contraction).
#include <stdarg.h>
#include "common.cuh"
#include "host_common.h"

namespace pic {

static thread_local char g_err[512] = "ok";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int device_sm_count() {
    static int sm = 0;
    if (!sm) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
        if (sm <= 0) sm = 148;
    }
    return sm;
}
int max_optin_smem() {
    static int v = 0;
    if (!v) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (v <= 0) v = 48 * 1024;
    }
    return v;
}

// ------------------------------------------------------------------ smoothing
__global__ void smooth_k(const double* __restrict__ F, double* __restrict__ out, int n, int variant) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int ip = (i + 1 == n) ? 0 : i + 1;
        int im = (i == 0) ? n - 1 : i - 1;
        // (np.roll(F,-1) + 2.0*F + np.roll(F,1)) * 0.25   [== /4.0 bit for bit]
        double v = ((F[ip] + 2.0 * F[i]) + F[im]) * 0.25;
        if (variant == 1 && (i == 0 || i == n - 1)) v = F[i];
        out[i] = v;
    }
}

// ------------------------------------------------------------------ differences
__global__ void differentiate_k(const double* __restrict__ F, double* __restrict__ out, int n, double dx,
                                int variant) {
    double idx_2 = 0.5 / dx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double v;
        if (variant == 0) {                       // pypic.py:204-210
            int ip = (i + 1 == n) ? 0 : i + 1;
            int im = (i == 0) ? n - 1 : i - 1;
            v = (F[ip] - F[im]) * idx_2;
        } else if (variant == 2) {                // PIC_L.py:238-243 (n = Ng+1 nodes)
            if (i == 0) v = -(F[1] - F[n - 1]) / dx * 0.5;
            else if (i == n - 1) v = -(F[0] - F[n - 2]) / dx * 0.5;
            else v = -(F[i + 1] - F[i - 1]) / dx * 0.5;
        } else {                                  // PIC_L_DD.py:195-200 / pygcpic.py:932-936
            if (i == 0) v = -(F[1] - F[0]) / dx;
            else if (i == n - 1) v = -(F[n - 1] - F[n - 2]) / dx;
            else if (variant == 1) v = -(F[i + 1] - F[i - 1]) / dx * 0.5;
            else v = -(F[i + 1] - F[i - 1]) / dx / 2.;
        }
        out[i] = v;
    }
}

// ------------------------------------------------------------------ cumulative trapezoid
// out[i] = -sum_{k<i} dx*(F[k+1]+F[k])/2 ; one CTA, chunked inclusive scan with carry.
__global__ void integrate_field_k(const double* __restrict__ F, double* __restrict__ out, int n, double dx,
                                  int subtract_max) {
    __shared__ double wsum[32];
    __shared__ double scratch[33];
    __shared__ double carry_s;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) { carry_s = 0.0; out[0] = -0.0 * 0.0; }
    __syncthreads();
    double mx = (threadIdx.x == 0) ? 0.0 : -INFINITY;   // out[0] = -0.0 -> compares as 0
    for (int base = 1; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        double t = (i < n) ? (dx * (F[i] + F[i - 1])) / 2.0 : 0.0;
        // warp inclusive scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        if (lane == 31) wsum[w] = t;
        __syncthreads();
        if (w == 0) {
            double s = lane < nw ? wsum[lane] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            wsum[lane] = s;   // inclusive warp totals
        }
        __syncthreads();
        double pre = carry_s + (w > 0 ? wsum[w - 1] : 0.0);
        double val = -(pre + t);
        if (i < n) { out[i] = val; mx = fmax(mx, val); }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = pre + t;
        __syncthreads();
    }
    if (subtract_max) {
        double m = block_reduce<1>(mx, scratch);
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = out[i] - m;
    }
}

__global__ void shift_extreme_k(const double* __restrict__ F, double* __restrict__ out, int n, int mode) {
    __shared__ double scratch[33];
    double m = mode == 0 ? -INFINITY : INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = mode == 0 ? fmax(m, F[i]) : fmin(m, F[i]);
    m = mode == 0 ? block_reduce<1>(m, scratch) : block_reduce<2>(m, scratch);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = F[i] - m;
}

// ------------------------------------------------------------------ PCR in shared memory
// In-place parallel cyclic reduction on n rows held in shared memory (a,b,c,d).
// Every thread owns rows tid, tid+T, ... (<= PCR_RPT rows), stages the new coefficients
// in registers, then all threads write back: two barriers per level, log2(n) levels.
#define PCR_THREADS 1024
#define PCR_RPT ((PIC_PCR_SMEM_MAX + PCR_THREADS - 1) / PCR_THREADS)

__device__ void pcr_smem(double* a, double* b, double* c, double* d, int n) {
    for (int s = 1; s < n; s <<= 1) {
        double na[PCR_RPT], nb[PCR_RPT], nc[PCR_RPT], nd[PCR_RPT];
#pragma unroll
        for (int r = 0; r < PCR_RPT; ++r) {
            int i = threadIdx.x + r * PCR_THREADS;
            if (i < n) {
                int im = i - s, ip = i + s;
                double ai = a[i], bi = b[i], ci = c[i], di = d[i];
                double k1 = 0.0, k2 = 0.0, am = 0.0, cm = 0.0, dm = 0.0, ap = 0.0, cp = 0.0, dp = 0.0;
                if (im >= 0) { k1 = ai / b[im]; am = a[im]; cm = c[im]; dm = d[im]; }
                if (ip < n) { k2 = ci / b[ip]; ap = a[ip]; cp = c[ip]; dp = d[ip]; }
                na[r] = -am * k1;
                nb[r] = bi - cm * k1 - ap * k2;
                nc[r] = -cp * k2;
                nd[r] = di - dm * k1 - dp * k2;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PCR_RPT; ++r) {
            int i = threadIdx.x + r * PCR_THREADS;
            if (i < n) { a[i] = na[r]; b[i] = nb[r]; c[i] = nc[r]; d[i] = nd[r]; }
        }
        __syncthreads();
    }
    // rows are decoupled: x = d/b, left in d
    for (int i = threadIdx.x; i < n; i += PCR_THREADS) d[i] = d[i] / b[i];
    __syncthreads();
}

__global__ void __launch_bounds__(PCR_THREADS) tridiag_pcr_smem_k(const double* __restrict__ a,
                                                                   const double* __restrict__ b,
                                                                   const double* __restrict__ c,
                                                                   const double* __restrict__ d,
                                                                   double* __restrict__ x, int n) {
    extern __shared__ double sm[];
    double *sa = sm, *sb = sm + n, *sc = sm + 2 * n, *sd = sm + 3 * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        sa[i] = (i == 0) ? 0.0 : a[i];
        sb[i] = b[i];
        sc[i] = (i == n - 1) ? 0.0 : c[i];
        sd[i] = d[i];
    }
    __syncthreads();
    pcr_smem(sa, sb, sc, sd, n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = sd[i];
}

// global-memory PCR level (n > PIC_PCR_SMEM_MAX): ping-pong between two coefficient sets
__global__ void pcr_level_k(const double* __restrict__ a, const double* __restrict__ b,
                            const double* __restrict__ c, const double* __restrict__ d, double* __restrict__ oa,
                            double* __restrict__ ob, double* __restrict__ oc, double* __restrict__ od, int n,
                            int s, int first) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int im = i - s, ip = i + s;
        double ai = (first && i == 0) ? 0.0 : a[i], bi = b[i], ci = (first && i == n - 1) ? 0.0 : c[i], di = d[i];
        double k1 = 0.0, k2 = 0.0, am = 0.0, cm = 0.0, dm = 0.0, ap = 0.0, cp = 0.0, dp = 0.0;
        if (im >= 0) {
            k1 = ai / b[im];
            am = (first && im == 0) ? 0.0 : a[im];
            cm = c[im];
            dm = d[im];
        }
        if (ip < n) {
            k2 = ci / b[ip];
            ap = a[ip];
            cp = (first && ip == n - 1) ? 0.0 : c[ip];
            dp = d[ip];
        }
        oa[i] = -am * k1;
        ob[i] = bi - cm * k1 - ap * k2;
        oc[i] = -cp * k2;
        od[i] = di - dm * k1 - dp * k2;
    }
}
__global__ void pcr_final_k(const double* __restrict__ b, const double* __restrict__ d, double* __restrict__ x,
                            int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = d[i] / b[i];
}

static int tridiag_launch(const double* a, const double* b, const double* c, const double* d, double* x, int n,
                          double* work, cudaStream_t st) {
    if (n <= PIC_PCR_SMEM_MAX) {
        size_t smem = (size_t)4 * n * sizeof(double);
        PIC_CHECK_CUDA(cudaFuncSetAttribute(tridiag_pcr_smem_k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        tridiag_pcr_smem_k<<<1, PCR_THREADS, smem, st>>>(a, b, c, d, x, n);
        PIC_CHECK_LAUNCH();
        return PIC_OK;
    }
    PIC_REQUIRE(work != nullptr, "tridiagonal solve with n > PIC_PCR_SMEM_MAX needs a work buffer of 8*n doubles");
    double* set[2][4];
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 4; ++j) set[k][j] = work + ((size_t)k * 4 + j) * n;
    int grid = grid_for(n, 256, 8);
    const double *ca = a, *cb = b, *cc = c, *cd = d;
    int cur = 0, first = 1;
    for (int s = 1; s < n; s <<= 1) {
        pcr_level_k<<<grid, 256, 0, st>>>(ca, cb, cc, cd, set[cur][0], set[cur][1], set[cur][2], set[cur][3], n, s,
                                          first);
        PIC_CHECK_LAUNCH();
        ca = set[cur][0]; cb = set[cur][1]; cc = set[cur][2]; cd = set[cur][3];
        cur ^= 1;
        first = 0;
    }
    pcr_final_k<<<grid, 256, 0, st>>>(cb, cd, x, n);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

// ------------------------------------------------------------------ Poisson solves
// builds the gauge-fixed periodic system (n-1 unknowns) or the Dirichlet system of
// Grid.solve_for_phi_dirichlet in global work arrays; mode 0 periodic, 1 dirichlet
__global__ void poisson_build_k(const double* __restrict__ rho, double* __restrict__ a, double* __restrict__ b,
                                double* __restrict__ c, double* __restrict__ d, int n, double dx, int mode) {
    __shared__ double scratch[33];
    double dx2 = dx * dx;
    if (mode == 0) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s += rho[i];
        s = block_reduce<0>(s, scratch);
        double c0 = -(s / (double)n) / PIC_EPS0;      // -np.average(rho)/epsilon0
        for (int i = threadIdx.x; i < n - 1; i += blockDim.x) {
            double c2 = rho[i] / PIC_EPS0;
            a[i] = 1.0; b[i] = -2.0; c[i] = 1.0;
            d[i] = -dx2 * c0 - dx2 * c2;
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            bool edge = (i == 0 || i == n - 1);
            a[i] = edge ? 0.0 : 1.0;
            c[i] = edge ? 0.0 : 1.0;
            b[i] = edge ? 1.0 : -2.0;
            d[i] = rho[i];
        }
    }
}
// mode 0: phi[n-1]=0, optional -max ; mode 1: phi = -x*dx2 then -min
__global__ void poisson_finish_k(const double* __restrict__ x, double* __restrict__ phi, int n, double dx,
                                 int mode, int subtract) {
    __shared__ double scratch[33];
    double dx2 = dx * dx;
    double m = mode == 0 ? -INFINITY : INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v;
        if (mode == 0) v = (i == n - 1) ? 0.0 : x[i];
        else v = -x[i] * dx2;
        phi[i] = v;
        m = mode == 0 ? fmax(m, v) : fmin(m, v);
    }
    m = mode == 0 ? block_reduce<1>(m, scratch) : block_reduce<2>(m, scratch);
    if (subtract)
        for (int i = threadIdx.x; i < n; i += blockDim.x) phi[i] = phi[i] - m;
}

// ------------------------------------------------------------------ Newton-Boltzmann, one launch
__global__ void __launch_bounds__(PCR_THREADS) newton_boltzmann_k(const double* __restrict__ src,
                                                                  double* __restrict__ phi, int n, double dx,
                                                                  double n0, double Te, int bc, double tol,
                                                                  int iter_max, int* __restrict__ iters_out) {
    extern __shared__ double sm[];
    __shared__ double scratch[33];
    double *sa = sm, *sb = sm + n, *sc = sm + 2 * n, *sd = sm + 3 * n;
    const double dx2 = dx * dx;
    const double c0 = PIC_E * n0 / PIC_EPS0;
    const double c1 = PIC_E / 1.38E-23 / Te;
    if (bc == 0)
        for (int i = threadIdx.x; i < n; i += blockDim.x) phi[i] = 0.0;   // cold start, pygcpic.py:1026
    __syncthreads();
    double residual = 1.0;
    int it = 0;
    while (residual > tol && it < iter_max) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double p = phi[i];
            double ex = exp(c1 * p);
            double c2 = (bc == 0) ? src[i] / PIC_EPS0 : PIC_E * src[i] / PIC_EPS0;
            double Aphi;
            if (i == 0) Aphi = p;
            else if (i == n - 1) Aphi = (bc == 0) ? p : (3.0 * p - 4.0 * phi[n - 2] + phi[n - 3]);
            else Aphi = phi[i - 1] - 2.0 * p + phi[i + 1];
            double F = Aphi - dx2 * c0 * ex + dx2 * c2;
            double D = -dx2 * c0 * c1 * ex;
            double av = 1.0, bv = -2.0 + D, cv = 1.0;
            if (i == 0) {
                F = (bc == 0) ? 0.0 : p;
                bv = 1.0 + (-dx2 * c0 * c1);
                av = 0.0; cv = 0.0;
            } else if (i == n - 1) {
                F = 0.0;
                if (bc == 0) { bv = 1.0 + (-dx2 * c0 * c1); av = 0.0; cv = 0.0; }
                else { bv = 3.0; av = -4.0; cv = 0.0; }   // folded below
            }
            sa[i] = av; sb[i] = bv; sc[i] = cv; sd[i] = F;
        }
        __syncthreads();
        if (bc == 1 && threadIdx.x == 0) {
            // last row [1,-4,3] minus row n-2 ([1, b, 1]) -> [0, -4-b, 2]; rhs F[n-1]-F[n-2]
            sa[n - 1] = -4.0 - sb[n - 2];
            sb[n - 1] = 3.0 - 1.0;
            sd[n - 1] = sd[n - 1] - sd[n - 2];
        }
        __syncthreads();
        pcr_smem(sa, sb, sc, sd, n);
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double dp = sd[i];
            phi[i] = phi[i] - dp;
            s += dp * dp;
        }
        s = block_reduce<0>(s, scratch);
        residual = (bc == 0) ? s : sqrt(s);
        ++it;
        __syncthreads();
    }
    double m = INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmin(m, phi[i]);
    m = block_reduce<2>(m, scratch);
    for (int i = threadIdx.x; i < n; i += blockDim.x) phi[i] = phi[i] - m;
    if (threadIdx.x == 0 && iters_out) *iters_out = it;
}


// ------------------------------------------------------------------ PIC_L / PIC_L_DD Boltzmann-Newton solves
// PIC_L.solvePoisson :146-177 = PIC_L_DD.solvePoisson :116-147 (periodic == 0) and
// PIC_L.solvePoissonPeriodic :179-206 = PIC_L_DD.solvePoissonPeriodic :149-176 (periodic == 1), whole
// Newton loop in one launch.  As written there: c0 = rho[mid]/eps0 (mid = n/2), c1 = e/kBT,
// F = A phi - dx^2 c0 exp(c1 (phi - phi[mid])) + dx^2 rho/eps0, J = A + diag(-dx^2 c0 c1 exp(..)),
// dphi = inv(J) F, loop `while resid > tol and k <= maxiter` with resid = |dphi|_2.
//  bounded: A = laplacian1D (first row e_0, LAST ROW [.., 1, 1, -2]); F[0] = phi[0], F[-1] = phi[-1];
//           J[0,0] = 1 - dx^2 c0 c1, J[-1,-1] = -2 - dx^2 c0 c1.  The three-entry last row is reduced with
//           row n-2 before the tridiagonal (PCR) solve.
//  periodic: A cyclic [1,-2,1]; the cyclic system is solved by Sherman-Morrison (two PCR solves).
__global__ void __launch_bounds__(PCR_THREADS) newton_boltzmann_l_k(const double* __restrict__ rho,
                                                                    double* __restrict__ phi, int n, double dx,
                                                                    double kBT, double tol, int maxiter, int periodic,
                                                                    double* __restrict__ work,
                                                                    int* __restrict__ iters_out) {
    extern __shared__ double sm[];
    __shared__ double scratch[33];
    __shared__ double s_bc[4];
    double *sa = sm, *sb = sm + n, *sc = sm + 2 * n, *sd = sm + 3 * n;
    const int mid = n / 2;
    const double dx2 = dx * dx;
    const double c0 = rho[mid] / PIC_EPS0;
    const double c1 = PIC_E / kBT;
    double resid = 1.0;
    int k = 0;
    while (resid > tol && k <= maxiter) {
        const double pm = phi[mid];
        // pass 0: the Newton system; periodic pass 1: the same matrix with the Sherman-Morrison vector u as rhs
        for (int pass = 0; pass < (periodic ? 2 : 1); ++pass) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const double p = phi[i];
                const double ex = exp(c1 * (p - pm));
                const double D = -dx2 * c0 * c1 * ex;
                double av = 1.0, bv = -2.0 + D, cv = 1.0, F;
                if (periodic) {
                    const double pl = phi[i == 0 ? n - 1 : i - 1], pr = phi[i == n - 1 ? 0 : i + 1];
                    F = (pl - 2.0 * p + pr) - dx2 * c0 * ex + dx2 * (rho[i] / PIC_EPS0);
                    // J = T + u v^T, gamma = -b0: T has b0 - gamma and b_{n-1} - 1/gamma on the corners' diagonal
                    const double gamma = -(-2.0 + (-dx2 * c0 * c1 * exp(c1 * (phi[0] - pm))));
                    if (i == 0) { bv = bv - gamma; av = 0.0; }
                    if (i == n - 1) { bv = bv - 1.0 / gamma; cv = 0.0; }
                    if (pass == 1) F = (i == 0) ? gamma : (i == n - 1 ? 1.0 : 0.0);
                    if (i == 0 && pass == 0) s_bc[0] = gamma;
                } else {
                    F = (i == 0 || i == n - 1) ? p : (phi[i - 1] - 2.0 * p + phi[i + 1]) - dx2 * c0 * ex + dx2 * (rho[i] / PIC_EPS0);
                    if (i == 0) { bv = 1.0 + (-dx2 * c0 * c1); av = 0.0; cv = 0.0; }
                    if (i == n - 1) { bv = -2.0 + (-dx2 * c0 * c1); av = 1.0; cv = 0.0; }     // + the entry at n-3, folded below
                }
                sa[i] = av; sb[i] = bv; sc[i] = cv; sd[i] = F;
            }
            __syncthreads();
            if (!periodic && threadIdx.x == 0) {
                // last row [1, 1, b] minus row n-2 ([1, b', 1]) -> [0, 1 - b', b - 1], rhs F[n-1] - F[n-2]
                sa[n - 1] = 1.0 - sb[n - 2];
                sb[n - 1] = sb[n - 1] - 1.0;
                sd[n - 1] = sd[n - 1] - sd[n - 2];
            }
            __syncthreads();
            pcr_smem(sa, sb, sc, sd, n);
            if (periodic && pass == 0) {
                for (int i = threadIdx.x; i < n; i += blockDim.x) work[i] = sd[i];      // y
                __syncthreads();
            }
        }
        double fac = 0.0;
        if (periodic) {
            // x = y - z (v.y)/(1 + v.z), v = [1, 0, ..., 0, 1/gamma]
            if (threadIdx.x == 0) {
                const double g = s_bc[0];
                const double vy = work[0] + work[n - 1] / g, vz = sd[0] + sd[n - 1] / g;
                s_bc[1] = vy / (1.0 + vz);
            }
            __syncthreads();
            fac = s_bc[1];
        }
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double dp = periodic ? work[i] - fac * sd[i] : sd[i];
            phi[i] = phi[i] - dp;
            s += dp * dp;
        }
        s = block_reduce<0>(s, scratch);
        resid = sqrt(s);
        ++k;
        __syncthreads();
    }
    if (threadIdx.x == 0 && iters_out) *iters_out = k;
}

}  // namespace pic

using namespace pic;

extern "C" {

const char* pic_last_error(void) { return g_err; }
int pic_version(void) { return 100; }

int pic_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
    int dev = 0, cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        set_error("no CUDA device visible");
        return PIC_ERR_NODEVICE;
    }
    PIC_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp pr;
    PIC_CHECK_CUDA(cudaGetDeviceProperties(&pr, dev));
    if (sm_count) *sm_count = pr.multiProcessorCount;
    if (cc_major) *cc_major = pr.major;
    if (cc_minor) *cc_minor = pr.minor;
    if (name && name_len > 0) {
        strncpy(name, pr.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return PIC_OK;
}

int pic_dev_smooth(const double* F, double* out, int n, int variant, void* stream) {
    PIC_REQUIRE(F && out && n >= 1 && F != out, "smooth: null/aliased pointer or n<1");
    smooth_k<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(F, out, n, variant);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_differentiate(const double* F, double* out, int n, double dx, int variant, void* stream) {
    PIC_REQUIRE(F && out && n >= 3 && F != out && variant >= 0 && variant <= 3, "differentiate: bad argument");
    differentiate_k<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(F, out, n, dx, variant);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_integrate_field(const double* F, double* out, int n, double dx, int subtract_max, void* stream) {
    PIC_REQUIRE(F && out && n >= 1 && F != out, "integrate_field: bad argument");
    integrate_field_k<<<1, 1024, 0, (cudaStream_t)stream>>>(F, out, n, dx, subtract_max);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_shift_extreme(const double* F, double* out, int n, int mode, void* stream) {
    PIC_REQUIRE(F && out && n >= 1, "shift_extreme: bad argument");
    shift_extreme_k<<<1, 1024, 0, (cudaStream_t)stream>>>(F, out, n, mode);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_tridiag_pcr(const double* a, const double* b, const double* c, const double* d, double* x, int n,
                        double* work, void* stream) {
    PIC_REQUIRE(a && b && c && d && x && n >= 1, "tridiag_pcr: bad argument");
    return tridiag_launch(a, b, c, d, x, n, work, (cudaStream_t)stream);
}

int pic_dev_poisson_periodic(const double* rho, double* phi, int n, double dx, int subtract_max, double* work,
                             void* stream) {
    PIC_REQUIRE(rho && phi && work && n >= 3, "poisson_periodic: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    double *a = work, *b = work + n, *c = work + 2 * (size_t)n, *d = work + 3 * (size_t)n, *x = work + 4 * (size_t)n;
    poisson_build_k<<<1, 1024, 0, st>>>(rho, a, b, c, d, n, dx, 0);
    PIC_CHECK_LAUNCH();
    int rc = tridiag_launch(a, b, c, d, x, n - 1, work + 5 * (size_t)n, st);   // needs 8*(n-1) more when large
    if (rc) return rc;
    poisson_finish_k<<<1, 1024, 0, st>>>(x, phi, n, dx, 0, subtract_max);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_poisson_dirichlet(const double* rho, double* phi, int n, double dx, double* work, void* stream) {
    PIC_REQUIRE(rho && phi && work && n >= 3, "poisson_dirichlet: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    double *a = work, *b = work + n, *c = work + 2 * (size_t)n, *d = work + 3 * (size_t)n, *x = work + 4 * (size_t)n;
    poisson_build_k<<<1, 1024, 0, st>>>(rho, a, b, c, d, n, dx, 1);
    PIC_CHECK_LAUNCH();
    int rc = tridiag_launch(a, b, c, d, x, n, work + 5 * (size_t)n, st);
    if (rc) return rc;
    poisson_finish_k<<<1, 1024, 0, st>>>(x, phi, n, dx, 1, 1);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_newton_boltzmann(const double* src, double* phi, int n, double dx, double n0, double Te, int bc,
                             double tol, int iter_max, int* iters_out, void* stream) {
    PIC_REQUIRE(src && phi && n >= 4 && (bc == 0 || bc == 1), "newton_boltzmann: bad argument");
    PIC_REQUIRE(n <= PIC_PCR_SMEM_MAX, "newton_boltzmann: n exceeds PIC_PCR_SMEM_MAX");
    size_t smem = (size_t)4 * n * sizeof(double);
    PIC_CHECK_CUDA(cudaFuncSetAttribute(newton_boltzmann_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    newton_boltzmann_k<<<1, PCR_THREADS, smem, (cudaStream_t)stream>>>(src, phi, n, dx, n0, Te, bc, tol, iter_max,
                                                                       iters_out);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_newton_boltzmann_l(const double* rho, double* phi, int n, double dx, double kBT, double tol, int maxiter,
                               int periodic, double* work, int* iters_out, void* stream) {
    PIC_REQUIRE(rho && phi && n >= 4 && kBT > 0 && (periodic == 0 || periodic == 1), "newton_boltzmann_l: bad argument");
    PIC_REQUIRE(n <= PIC_PCR_SMEM_MAX, "newton_boltzmann_l: n exceeds PIC_PCR_SMEM_MAX");
    PIC_REQUIRE(!periodic || work, "newton_boltzmann_l: the periodic solve needs a work buffer of n doubles");
    size_t smem = (size_t)4 * n * sizeof(double);
    PIC_CHECK_CUDA(cudaFuncSetAttribute(newton_boltzmann_l_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    newton_boltzmann_l_k<<<1, PCR_THREADS, smem, (cudaStream_t)stream>>>(rho, phi, n, dx, kBT, tol, maxiter, periodic, work,
                                                                        iters_out);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

}  // extern "C"
