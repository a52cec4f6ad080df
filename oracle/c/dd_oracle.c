/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the Picard timestep of
 * PIC_L_DD.main_i (reference PIC_L_DD.py:452-545 with interpolateField :32-39 and
 * weightCurrents :41-68).  Used (a) as a scalable checker for sizes the NumPy oracle
 * is too slow for and (b) as the CPU baseline that bench.py times (kind "port").
 * Never linked into or called by the product path.
 *
 * Parity: pinned against oracle/np_oracle.py (itself pinned against the reference's
 * golden vectors) in tests/test_oracle.py::test_c_oracle_*.
 * Compile with -ffp-contract=off so that no a*b+c is fused (NumPy does not fuse).
 * With nthreads==1 the deposit is the reference's serial particle-order loop; with
 * nthreads>1 (OpenMP) each thread owns a private grid that is reduced at the end
 * (timing baseline; sums agree to round-off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPS0 8.854E-12

static inline double pymod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) { if ((b < 0.0) != (r < 0.0)) r += b; } else r = copysign(0.0, b);
    return r;
}

/* one Picard iteration over particles [lo,hi): gather, push, absorb, deposit */
static void iter_range(long lo, long hi, long n_split, int Ng, double dx, double dt, double L, double p2c,
                       const double* q, const double* m, const double* x0, const double* u0, double* x1,
                       double* u1, double* xs, double* active, const double* Es, double* jh, double* j1,
                       int first) {
    const double idx = 1. / dx;
    for (long i = lo; i < hi; ++i) {
        int sp = i >= n_split;
        if (active[i] != 1.0) {
            if (active[i] == -1.0) { jh[0] += dx * q[sp] * p2c / dt; j1[0] += dx * q[sp] * p2c / dt; }
            else { jh[Ng - 1] += -dx * q[sp] * p2c / dt; j1[Ng - 1] += -dx * q[sp] * p2c / dt; }
            x1[i] = 0.0; u1[i] = 0.0;
            continue;
        }
        double X0 = x0[i], U0 = u0[i];
        double s = first ? X0 : xs[i];
        int ind = (int)floor(s / dx);
        double wR = pymod(s, dx) / dx, wL = 1. - wR;
        double Ei = wL * Es[ind] + wR * Es[ind + 1];
        double qm = q[sp] / m[sp];
        double X1 = X0 + dt * U0 + dt * dt * qm * Ei * 0.5;
        double U1 = U0 + dt * qm * Ei;
        double XH = (X0 + X1) * 0.5, UH = (U0 + U1) * 0.5;
        x1[i] = X1; u1[i] = U1; xs[i] = XH;
        if (X0 >= L || XH >= L || X1 >= L) {
            active[i] = 0.0;
            jh[Ng - 1] += -dx * q[sp] * p2c / dt; j1[Ng - 1] += -dx * q[sp] * p2c / dt;
            continue;
        }
        if (X0 <= 0.0 || XH <= 0.0 || X1 <= 0.0) {
            active[i] = -1.0;
            jh[0] += dx * q[sp] * p2c / dt; j1[0] += dx * q[sp] * p2c / dt;
            continue;
        }
        int ih = (int)floor(XH / dx);
        double hR = pymod(XH, dx) / dx, hL = 1. - hR;
        jh[ih] += q[sp] * UH * p2c * hL * idx;
        jh[ih + 1] += q[sp] * UH * p2c * hR * idx;
        int i1 = (int)floor(X1 / dx);
        double fR = pymod(X1, dx) / dx, fL = 1. - fR;
        j1[i1] += q[sp] * U1 * p2c * fL * idx;
        j1[i1 + 1] += q[sp] * U1 * p2c * fR * idx;
    }
}

/* returns the iteration count; *resid gets the last residual.
 * NOTE the wall terms: the reference adds them in particle order inside weightCurrents
 * for EVERY inactive particle on every call; here particles absorbed in this iteration
 * add theirs at their own position in the loop, which is the same serial order. */
int dd_picard_step_c(long N, long n_split, int Ng, double dx, double dt, double L, double p2c, const double* q,
                     const double* m, const double* x0, const double* u0, double* active, const double* E0,
                     double tol, int maxiter, double* x1, double* u1, double* E1, double* j1out, double* resid,
                     int nthreads) {
    double* xs = (double*)malloc(sizeof(double) * (size_t)N);
    double* Es = (double*)malloc(sizeof(double) * Ng);
    double* jh = (double*)malloc(sizeof(double) * Ng);
    double* j1 = (double*)malloc(sizeof(double) * Ng);
    memcpy(Es, E0, sizeof(double) * Ng);
    memcpy(E1, E0, sizeof(double) * Ng);
    double r = 1.0;
    int k = 0;
#ifndef _OPENMP
    nthreads = 1;
#endif
    if (nthreads < 1) nthreads = 1;
    double* priv = nthreads > 1 ? (double*)malloc(sizeof(double) * 2 * Ng * (size_t)nthreads) : NULL;
    while (r > tol && k < maxiter) {
        memset(jh, 0, sizeof(double) * Ng);
        memset(j1, 0, sizeof(double) * Ng);
        if (nthreads == 1) {
            iter_range(0, N, n_split, Ng, dx, dt, L, p2c, q, m, x0, u0, x1, u1, xs, active, Es, jh, j1, k == 0);
        } else {
#ifdef _OPENMP
            memset(priv, 0, sizeof(double) * 2 * Ng * (size_t)nthreads);
#pragma omp parallel num_threads(nthreads)
            {
                /* the runtime may grant fewer threads than requested: partition by what it granted */
                int t = omp_get_thread_num(), nt = omp_get_num_threads();
                long lo = (long)((double)N * t / nt), hi = t + 1 == nt ? N : (long)((double)N * (t + 1) / nt);
                iter_range(lo, hi, n_split, Ng, dx, dt, L, p2c, q, m, x0, u0, x1, u1, xs, active, Es,
                           priv + 2 * (size_t)Ng * t, priv + 2 * (size_t)Ng * t + Ng, k == 0);
            }
            for (int t = 0; t < nthreads; ++t)
                for (int g = 0; g < Ng; ++g) { jh[g] += priv[2 * (size_t)Ng * t + g]; j1[g] += priv[2 * (size_t)Ng * t + Ng + g]; }
#endif
        }
        jh[0] += jh[1]; jh[Ng - 1] += jh[Ng - 2];
        j1[0] += j1[1]; j1[Ng - 1] += j1[Ng - 2];
        double mean = 0.0;
        for (int g = 0; g < Ng; ++g) mean += jh[g];
        mean /= (double)Ng;
        double rr = 0.0;
        for (int g = 0; g < Ng; ++g) {
            double e1 = E0[g] + (dt / EPS0) * (mean - jh[g]);
            double eh = (e1 + E0[g]) * 0.5;
            double d = Es[g] - eh;
            rr += d * d;
            E1[g] = e1;
            Es[g] = eh;
        }
        r = sqrt(rr);
        ++k;
    }
    memcpy(j1out, j1, sizeof(double) * Ng);
    *resid = r;
    free(xs); free(Es); free(jh); free(j1);
    if (priv) free(priv);
    return k;
}

int dd_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* threads the runtime actually grants to a team of the requested size */
int dd_oracle_team_size(int nthreads) {
    int got = 1;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
#pragma omp single
        got = omp_get_num_threads();
    }
#endif
    return got;
}
