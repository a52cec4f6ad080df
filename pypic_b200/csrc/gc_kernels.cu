// pygcpic.py -- Particle/Grid hot path on a structure-of-arrays particle store:
// mirrored CIC gather, Boris 1D3V push, 6D<->guiding-centre transforms, GC RK4 push,
// Dirichlet wall absorption, (rho,n) deposit, Boltzmann reference-density update, the
// order-dependent reactivate-or-delete rule as a prefix scan, and warp-ballot stable
// stream compaction.
#include <stdlib.h>
#include "common.cuh"
#include "ring.cuh"
#include "host_common.h"

namespace pic {

struct GCK {
    long long N;
    int ng, flags;
    double dx, dt, length;
    double B[3], Eyz[2];
};
static GCK make_gck(const pic_gc_params* p) {
    GCK k;
    k.N = p->N; k.ng = p->ng; k.flags = p->flags; k.dx = p->dx; k.dt = p->dt; k.length = p->length;
    for (int i = 0; i < 3; ++i) k.B[i] = p->B[i];
    k.Eyz[0] = p->Eyz[0]; k.Eyz[1] = p->Eyz[1];
    return k;
}
struct R7 { double* r[7]; };

// pygcpic.py:344-347: the LEFT node is weighted with the fractional distance (mirrored)
// idx = RN(1/dx): the division-free lookup (cell_dd_fast) is bit-identical to cell_dd
__device__ __forceinline__ double gather_mirrored(const double* E, double x, double dx, double idx, int ng, int& bad) {
    Cell c = cell_dd_fast(x, dx, idx);
    if (c.iL < 0 || c.iL > ng - 2) { ++bad; c.iL = clampi(c.iL, 0, ng - 2); }
    double w_l = c.wR;            // (x%dx)/dx
    double w_r = 1.0 - w_l;
    return E[c.iL] * w_l + E[c.iL + 1] * w_r;
}

__global__ void gc_interpolate_k(const double* __restrict__ E, const double* __restrict__ x, double* __restrict__ out,
                                 long long N, int ng, double dx, int* __restrict__ range_err) {
    int bad = 0;
    const double idx = 1.0 / dx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        out[i] = gather_mirrored(E, x[i], dx, idx, ng, bad);
    if (bad && range_err) atomicAdd(range_err, bad);
}

// pygcpic.py:871-883
__global__ void gc_weight_k(const double* __restrict__ x, const double* __restrict__ cs, const double* __restrict__ p2c,
                            const int8_t* __restrict__ active, double* __restrict__ rho, double* __restrict__ n,
                            long long N, int ng, double dx, int* __restrict__ range_err) {
    extern __shared__ double sm[];
    double *sr = sm, *sn = sm + ng;
    for (int i = threadIdx.x; i < 2 * ng; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    int bad = 0;
    const double idx = 1.0 / dx;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nIter = (N + stride - 1) / stride;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long it = 0; it < nIter; ++it, i += stride) {
        const bool valid = i < N && active[i] == 1;
        int iL = -1;
        double a0 = 0., a1 = 0., b0 = 0., b1 = 0.;
        if (valid) {
            Cell c = cell_dd_fast(x[i], dx, idx);      // bit-identical to cell_dd, no IEEE division
            if (c.iL < 0 || c.iL > ng - 2) { ++bad; c.iL = clampi(c.iL, 0, ng - 2); }
            double pc = p2c[i];
            double qr = div_const(cs[i] * PIC_E * pc, dx, idx);      // charge_state*e*p2c/dx
            double nr = div_const(pc, dx, idx);
            iL = c.iL;
            a0 = qr * c.wL; a1 = qr * c.wR; b0 = nr * c.wL; b1 = nr * c.wR;
        }
        // warp pre-reduction when the whole warp sits in one cell (store sorted by cell): fp64
        // shared-memory atomics are CAS loops, 32-way same-address conflicts serialise badly
        const int k0 = __shfl_sync(0xffffffffu, iL, 0);
        if (__all_sync(0xffffffffu, iL == k0)) {
            if (k0 < 0) continue;
            a0 = warp_sum(a0); a1 = warp_sum(a1); b0 = warp_sum(b0); b1 = warp_sum(b1);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&sr[k0], a0); atomicAdd(&sr[k0 + 1], a1); atomicAdd(&sn[k0], b0); atomicAdd(&sn[k0 + 1], b1);
            }
        } else if (valid) {
            atomicAdd(&sr[iL], a0);
            atomicAdd(&sr[iL + 1], a1);
            atomicAdd(&sn[iL], b0);
            atomicAdd(&sn[iL + 1], b1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ng; i += blockDim.x) {
        if (sr[i] != 0.0) atomicAdd(&rho[i], sr[i]);
        if (sn[i] != 0.0) atomicAdd(&n[i], sn[i]);
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

// fused gather -> Boris (pygcpic.py:478-506) -> Dirichlet BC (:685-687)
__global__ void __launch_bounds__(256) gc_push_boris_k(GCK k, R7 r, const double* __restrict__ cs,
                                                       const double* __restrict__ m, int8_t* __restrict__ active,
                                                       int8_t* __restrict__ at_wall, int8_t* __restrict__ hit_flag,
                                                       const double* __restrict__ Egrid,
                                                       long long* __restrict__ hit_count, int* __restrict__ range_err) {
    extern __shared__ double sE[];
    const double* E = Egrid;
    const bool pre = k.flags & 1;       // Egrid holds one already-gathered E_x per particle (Particle.push_6D alone)
    const bool nobc = k.flags & 2;      // no boundary test (Particle.push_6D alone)
    if ((k.flags & 4) && !pre) {   // field tile in shared memory
        for (int i = threadIdx.x; i < k.ng; i += blockDim.x) sE[i] = Egrid[i];
        __syncthreads();
        E = sE;
    }
    int bad = 0, hits = 0;
    const double idx = 1.0 / k.dx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        if (active[i] != 1) { if (hit_flag) hit_flag[i] = 0; continue; }
        double x = ld_stream(r.r[0] + i), y = ld_stream(r.r[1] + i), z = ld_stream(r.r[2] + i);
        double vx = ld_stream(r.r[3] + i), vy = ld_stream(r.r[4] + i), vz = ld_stream(r.r[5] + i);
        double Ex = pre ? Egrid[i] : gather_mirrored(E, x, k.dx, idx, k.ng, bad);
        double constant = 0.5 * k.dt * cs[i] * 1.602e-19 / m[i];
        vx += constant * Ex;
        double tx = constant * k.B[0], ty = constant * k.B[1], tz = constant * k.B[2];
        double t2 = tx * tx + ty * ty + tz * tz;
        double sx = 2. * tx / (1. + t2), sy = 2. * ty / (1. + t2), sz = 2. * tz / (1. + t2);
        double vfx = vx + vy * tz - vz * ty;
        double vfy = vy + vz * tx - vx * tz;
        double vfz = vz + vx * ty - vy * tx;
        vx += vfy * sz - vfz * sy;
        vy += vfz * sx - vfx * sz;
        vz += vfx * sy - vfy * sx;
        vx += constant * Ex;
        x += vx * k.dt; y += vy * k.dt; z += vz * k.dt;
        st_stream(r.r[0] + i, x); st_stream(r.r[1] + i, y); st_stream(r.r[2] + i, z);
        st_stream(r.r[3] + i, vx); st_stream(r.r[4] + i, vy); st_stream(r.r[5] + i, vz);
        r.r[6][i] = r.r[6][i] + k.dt;
        bool hit = !nobc && ((x < 0.0) || (x > k.length));
        if (hit) { active[i] = 0; at_wall[i] = 1; ++hits; }
        if (hit_flag) hit_flag[i] = hit ? 1 : 0;
    }
    if (hits && hit_count) atomicAdd((unsigned long long*)hit_count, (unsigned long long)hits);
    if (bad && range_err) atomicAdd(range_err, bad);
}

// =====================================================================================
// v2 fused kernel for a SPECIES-UNIFORM store (every particle has the same charge_state, m,
// p2c -- the host checks): gather (mirrored weights) + Boris 1D3V + Dirichlet walls + optional
// CIC deposit of the number density at the NEW position (the next step's D5 for the particles
// still active; rho = charge_state*e*n for a uniform store), in one pass.  Same design as
// dd_picard_iter_v6_k: persistent 512-thread CTA per SM, rows of 64 particles of all seven
// components staged by 1-D TMA bulk copies into a per-warp ring, two particles per lane,
// private shared-memory deposit windows.  Traffic: 7x8 B in + 7x8 B out + 1 B flag per particle.
#define G_T 512
// deposit window of G_W nodes per thread.  Hydrogen ions at the reference's resolution cross ~0.6 cells per
// step, so a 7-node window (the sheath kernel's) is left after 2-3 steps; 15 nodes hold 8 steps of drift:
// measured at 2e8 particles (profiles/r2_boris_window_width.txt) 5.1e10 particle-steps/s with a sort every
// 8 steps against 4.2e10 with 7 nodes and a sort every 4
#ifndef G_W
#define G_W 15
#endif
#define G_ROWS 16
#define G_CHUNK (G_T * 2 * G_ROWS)

// lean: the store does not track y, z and the per-particle clock (nothing on the path reads them: E
// depends on x only); the kernel then streams x,vx,vy,vz alone -- 64 B per particle-step, the row of
// SURVEY.md 8(d) -- and a particle that hits a wall gets its clock r[6] = tnow (the time of this push)
struct GUni { double cs, m, p2c; int lean; double tnow; };
struct GFastC {
    double dx, idx, dt, cE, tx, ty, tz, sx, sy, sz, nr;
    unsigned hi_lim;
};
struct GFastO { double x, y, z, vx, vy, vz, t, fL, fR; int cF; unsigned fr, ps; };

__device__ __forceinline__ void gc_fast(const GFastC& c, const double* __restrict__ sE, int ng, double x, double y,
                                        double z, double vx, double vy, double vz, double t, GFastO& o) {
    const double ts = x * c.idx, fs = floor(ts);
    const unsigned f0 = (unsigned)__double2hiint(ts - fs) - PIC_HI_G;
    const double rs = fma(-fs, c.dx, x);
    const int is = min(max((int)fs, 0), ng - 2);
    const double w_l = div_const(rs, c.dx, c.idx), w_r = 1.0 - w_l;       // pygcpic.py:344-346 (mirrored)
    const double Ex = sE[is] * w_l + sE[is + 1] * w_r;
    vx += c.cE * Ex;                                                      // :480
    const double vfx = vx + vy * c.tz - vz * c.ty;                        // :492-494
    const double vfy = vy + vz * c.tx - vx * c.tz;
    const double vfz = vz + vx * c.ty - vy * c.tx;
    vx += vfy * c.sz - vfz * c.sy;                                        // :496-498
    vy += vfz * c.sx - vfx * c.sz;
    vz += vfx * c.sy - vfy * c.sx;
    vx += c.cE * Ex;                                                      // :500
    o.x = x + vx * c.dt; o.y = y + vy * c.dt; o.z = z + vz * c.dt;        // :502-504
    o.vx = vx; o.vy = vy; o.vz = vz;
    o.t = t + c.dt;                                                       // :506
    o.ps = max((unsigned)__double2hiint(x) - 1u, (unsigned)__double2hiint(o.x) - 1u);
    const double tf = o.x * c.idx, ff = floor(tf);
    const unsigned f1 = (unsigned)__double2hiint(tf - ff) - PIC_HI_G;
    const double rf = fma(-ff, c.dx, o.x);
    o.cF = (int)ff;
    o.fr = max(f0, f1);
    o.fR = c.nr * (rf * c.idx); o.fL = c.nr - o.fR;                       // :880-883 up to re-association
}

// exact per-particle routine (the body of gc_push_boris_k + gc_weight_k's n deposit)
__device__ __noinline__ int gc_particle_exact(const GCK& k, const GUni& u, const R7& r, long long i, int act,
                                              const double* sE, int8_t* __restrict__ active,
                                              int8_t* __restrict__ at_wall, int8_t* __restrict__ hit_flag,
                                              double* __restrict__ n_acc, int* hits) {
    int bad = 0;
    if (act != 1) { if (hit_flag) hit_flag[i] = 0; return 0; }
    double x = r.r[0][i], y = u.lean ? 0.0 : r.r[1][i], z = u.lean ? 0.0 : r.r[2][i], vx = r.r[3][i], vy = r.r[4][i], vz = r.r[5][i];
    const double Ex = gather_mirrored(sE, x, k.dx, 1.0 / k.dx, k.ng, bad);
    const double constant = 0.5 * k.dt * u.cs * 1.602e-19 / u.m;
    vx += constant * Ex;
    const double tx = constant * k.B[0], ty = constant * k.B[1], tz = constant * k.B[2];
    const double t2 = tx * tx + ty * ty + tz * tz;
    const double sx = 2. * tx / (1. + t2), sy = 2. * ty / (1. + t2), sz = 2. * tz / (1. + t2);
    const double vfx = vx + vy * tz - vz * ty;
    const double vfy = vy + vz * tx - vx * tz;
    const double vfz = vz + vx * ty - vy * tx;
    vx += vfy * sz - vfz * sy;
    vy += vfz * sx - vfx * sz;
    vz += vfx * sy - vfy * sx;
    vx += constant * Ex;
    x += vx * k.dt; y += vy * k.dt; z += vz * k.dt;
    r.r[0][i] = x; r.r[3][i] = vx; r.r[4][i] = vy; r.r[5][i] = vz;
    const bool hit = (x < 0.0) || (x > k.length);                         // :685
    if (!u.lean) { r.r[1][i] = y; r.r[2][i] = z; r.r[6][i] = r.r[6][i] + k.dt; }
    else if (hit) r.r[6][i] = u.tnow;
    if (hit) { active[i] = 0; at_wall[i] = 1; ++*hits; }
    if (hit_flag) hit_flag[i] = hit ? 1 : 0;
    if (!hit && n_acc) {
        Cell c = cell_dd(x, k.dx);
        if (c.iL < 0 || c.iL > k.ng - 2) { ++bad; c.iL = clampi(c.iL, 0, k.ng - 2); }
        const double nr = u.p2c / k.dx;
        atomicAdd(&n_acc[c.iL], nr * c.wL);
        atomicAdd(&n_acc[c.iL + 1], nr * c.wR);
    }
    return bad;
}

__device__ __forceinline__ void gwin_add(double* myw, double* __restrict__ acc, int wb, int c, double vL, double vR) {
    const unsigned d = (unsigned)(c - wb);
    if (d <= (unsigned)(G_W - 2)) { double* p = myw + d * G_T; p[0] += vL; p[G_T] += vR; }
    else { atomicAdd(&acc[c], vL); atomicAdd(&acc[c + 1], vR); }
}

template <int NST, bool DEP, bool LEAN = false>
__global__ void __launch_bounds__(G_T, 1) gc_push_boris_v2_k(const __grid_constant__ GCK k, const GUni u, int nchunks,
                                                              const R7 r, int8_t* __restrict__ active,
                                                              int8_t* __restrict__ at_wall, int8_t* __restrict__ hit_flag,
                                                              const double* __restrict__ Egrid,
                                                              double* __restrict__ n_acc,
                                                              long long* __restrict__ hit_count,
                                                              int* __restrict__ range_err) {
    extern __shared__ __align__(128) double sm[];
    __shared__ int s_cnt[2];
    const int ng = k.ng;
    const int NP = (ng + 15) & ~15;
    double* sE = sm;
    double* win = sm + NP;                                   // [G_W][G_T]
    constexpr int SDP = LEAN ? 256 : 448;                    // doubles of particle data per stage: 4 or 7 arrays x 64 particles
    constexpr int SD = SDP + 16;                             // + the 64 activity flags of the row (padded to 128 B)
    double* ring = win + (DEP ? G_W * G_T : 0);              // [warp][stage][7 or 4 arrays][64] + flags
    unsigned long long* bars = (unsigned long long*)(ring + (G_T / 32) * NST * SD);
    for (int i = threadIdx.x; i < ng; i += G_T) sE[i] = Egrid[i];
    double* myw = win + threadIdx.x;
    if (DEP) {
#pragma unroll
        for (int n = 0; n < G_W; ++n) myw[n * G_T] = 0.0;
    }
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wbase = threadIdx.x & ~31;
    const int NOWIN = -0x40000000;
    const double* wring = ring + warp * (NST * SD);
    const uint32_t ring_s = smem_u32(wring);
    const uint32_t bar_s = smem_u32(bars + warp * NST);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    GFastC fc;
    fc.dx = k.dx; fc.idx = 1.0 / k.dx; fc.dt = k.dt;
    fc.cE = 0.5 * k.dt * u.cs * 1.602e-19 / u.m;                          // pygcpic.py:478
    fc.tx = fc.cE * k.B[0]; fc.ty = fc.cE * k.B[1]; fc.tz = fc.cE * k.B[2];
    {
        const double t2 = fc.tx * fc.tx + fc.ty * fc.ty + fc.tz * fc.tz;
        fc.sx = 2. * fc.tx / (1. + t2); fc.sy = 2. * fc.ty / (1. + t2); fc.sz = 2. * fc.tz / (1. + t2);
    }
    fc.nr = u.p2c / k.dx;
    fc.hi_lim = (unsigned)__double2hiint(k.length) - 1u;
    const int my_chunks = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long woff = (long long)warp * (64 * G_ROWS);
    const long long chunk_step = (long long)gridDim.x * G_CHUNK;
    auto issue = [&](long long base, int st) {
        if (elect_one()) {
            const uint32_t dst = ring_s + st * (SD * 8), bar = bar_s + 8 * st;
            mbar_expect_tx(bar, (uint32_t)(SDP * 8 + 64));
            // the row's activity flags ride in the same stage: a plain global load of them at the top of
            // every row would expose a full memory latency per row (they gate the fast path)
            bulk_g2s(dst + SDP * 8, active + base, 64, bar);
            if (LEAN) {
                bulk_g2s(dst, r.r[0] + base, 512, bar);
#pragma unroll
                for (int a = 3; a < 6; ++a) bulk_g2s(dst + 512 * (a - 2), r.r[a] + base, 512, bar);
            } else {
#pragma unroll
                for (int a = 0; a < 7; ++a) bulk_g2s(dst + 512 * a, r.r[a] + base, 512, bar);
            }
        }
    };
    long long cbase = (long long)blockIdx.x * G_CHUNK + woff;
    if (my_chunks > 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) issue(cbase + 64 * s, s);
    }
    int stage = 0, bad = 0, hits = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (int c = 0; c < my_chunks; ++c, cbase += chunk_step) {
        const bool more = c + 1 < my_chunks;
        int wb = NOWIN;
        long long ci = cbase + 2 * lane;
#pragma unroll 1
        for (int row = 0; row < G_ROWS; ++row, ci += 64) {
            mbar_wait(bar_s + 8 * stage, phase);
            const short fl = *(const short*)((const char*)(wring + stage * SD + SDP) + 2 * lane);
            const int acta = (int)(signed char)(fl & 0xff), actb = (int)(signed char)(fl >> 8);
            const double* sb = wring + stage * SD + 2 * lane;
            const double2 X = *(const double2*)sb;
            const double2 zero2 = make_double2(0., 0.);
            const double2 Y = LEAN ? zero2 : *(const double2*)(sb + 64), Z = LEAN ? zero2 : *(const double2*)(sb + 128);
            const double2 VX = *(const double2*)(sb + (LEAN ? 64 : 192)), VY = *(const double2*)(sb + (LEAN ? 128 : 256));
            const double2 VZ = *(const double2*)(sb + (LEAN ? 192 : 320));
            const double2 T = LEAN ? zero2 : *(const double2*)(sb + 384);
            const int st_cur = stage;
            if (++stage == NST) { stage = 0; phase ^= 1u; }
            GFastO a, b;
            gc_fast(fc, sE, ng, X.x, Y.x, Z.x, VX.x, VY.x, VZ.x, T.x, a);
            gc_fast(fc, sE, ng, X.y, Y.y, Z.y, VX.y, VY.y, VZ.y, T.y, b);
            const bool ra = (a.fr > PIC_HI_SPAN) | (a.ps >= fc.hi_lim) | (acta != 1);
            const bool rb = (b.fr > PIC_HI_SPAN) | (b.ps >= fc.hi_lim) | (actb != 1);
            if (DEP && row == 0) {
                int nok = __reduce_add_sync(full, (ra ? 0 : 1) + (rb ? 0 : 1));
                int sum = __reduce_add_sync(full, (ra ? 0 : a.cF) + (rb ? 0 : b.cF));
                wb = nok ? sum / nok - (G_W - 2) / 2 : NOWIN;
            }
            if (!(ra | rb)) {
                __stcs((double2*)(r.r[0] + ci), make_double2(a.x, b.x));
                if (!LEAN) {
                    __stcs((double2*)(r.r[1] + ci), make_double2(a.y, b.y));
                    __stcs((double2*)(r.r[2] + ci), make_double2(a.z, b.z));
                }
                __stcs((double2*)(r.r[3] + ci), make_double2(a.vx, b.vx));
                __stcs((double2*)(r.r[4] + ci), make_double2(a.vy, b.vy));
                __stcs((double2*)(r.r[5] + ci), make_double2(a.vz, b.vz));
                if (!LEAN) __stcs((double2*)(r.r[6] + ci), make_double2(a.t, b.t));
                if (DEP) { gwin_add(myw, n_acc, wb, a.cF, a.fL, a.fR); gwin_add(myw, n_acc, wb, b.cF, b.fL, b.fR); }
            } else {
                if (ra) bad += gc_particle_exact(k, u, r, ci, acta, sE, active, at_wall, hit_flag, DEP ? n_acc : nullptr, &hits);
                else {
                    r.r[0][ci] = a.x; r.r[3][ci] = a.vx; r.r[4][ci] = a.vy; r.r[5][ci] = a.vz;
                    if (!LEAN) { r.r[1][ci] = a.y; r.r[2][ci] = a.z; r.r[6][ci] = a.t; }
                    if (DEP) gwin_add(myw, n_acc, wb, a.cF, a.fL, a.fR);
                }
                if (rb) bad += gc_particle_exact(k, u, r, ci + 1, actb, sE, active, at_wall, hit_flag, DEP ? n_acc : nullptr, &hits);
                else {
                    r.r[0][ci + 1] = b.x; r.r[3][ci + 1] = b.vx; r.r[4][ci + 1] = b.vy; r.r[5][ci + 1] = b.vz;
                    if (!LEAN) { r.r[1][ci + 1] = b.y; r.r[2][ci + 1] = b.z; r.r[6][ci + 1] = b.t; }
                    if (DEP) gwin_add(myw, n_acc, wb, b.cF, b.fL, b.fR);
                }
            }
            __syncwarp();
            if (row < G_ROWS - NST) issue(cbase + 64 * (row + NST), st_cur);
            else if (more) issue(cbase + chunk_step + 64 * (row + NST - G_ROWS), st_cur);
        }
        __syncwarp();
        if (DEP && wb != NOWIN) {
            double s = 0.0;
            const int n = lane >> 1, half = lane & 1;
            if (n < G_W) {
                const double* col = win + n * G_T + wbase + half * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) s += col[(j + n) & 15];
            }
            s += __shfl_xor_sync(full, s, 1);
            if (n < G_W && half == 0) {
                const int node = wb + n;
                if (node >= 0 && node < ng && s != 0.0) atomicAdd(&n_acc[node], s);
            }
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < G_W; ++n2) myw[n2 * G_T] = 0.0;
            __syncwarp();
        }
    }
    // the N % G_CHUNK particles behind the last whole chunk: at most one per thread, exact routine
    for (long long i = (long long)nchunks * G_CHUNK + (long long)blockIdx.x * G_T + threadIdx.x; i < k.N; i += (long long)gridDim.x * G_T)
        bad += gc_particle_exact(k, u, r, i, active[i], sE, active, at_wall, hit_flag, DEP ? n_acc : nullptr, &hits);
    if (bad) atomicAdd(&s_cnt[0], bad);
    if (hits) atomicAdd(&s_cnt[1], hits);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0] && range_err) atomicAdd(range_err, s_cnt[0]);
        if (s_cnt[1] && hit_count) atomicAdd((unsigned long long*)hit_count, (unsigned long long)s_cnt[1]);
    }
}

// =====================================================================================
// The same fused step for a MIXED store (ions, neutrals, boron in several charge states in one
// list, as in the reference's particle loop pygcpic.py:1498-1549): charge_state, m and p2c are
// per-particle arrays and ride the TMA ring next to x, vx, vy, vz and the activity flags; the Boris
// constants are formed per particle with the IEEE operations of gc_push_boris_k (bit-identical
// results), and BOTH deposits of D5 are fused: n (p2c/dx) and rho (charge_state*e*p2c/dx) in two
// private 7-node windows per thread.  y, z and the clock (a full store) are only accumulated, so they
// bypass the ring: loaded at the top of a row, stored at its end.  Two ring stages.
#define GM_W 7
#define GM_NST 2
#define GM_SDP (7 * 64)
#define GM_SD (GM_SDP + 16)
struct GMix { const double* cs; const double* m; const double* p2c; int lean; double tnow; };
struct GMixO { double x, vx, vy, vz, nL, nR, qL, qR; int cF; unsigned fr, ps; };

__device__ __forceinline__ void gm_fast(const GCK& k, double idx, const double* __restrict__ sE, int ng, double x, double vx,
                                        double vy, double vz, double cs, double m, double pc, GMixO& o) {
    const double ts = x * idx, fs = floor(ts);
    const unsigned f0 = (unsigned)__double2hiint(ts - fs) - PIC_HI_G;
    const double rs = fma(-fs, k.dx, x);
    const int is = min(max((int)fs, 0), ng - 2);
    const double w_l = div_const(rs, k.dx, idx), w_r = 1.0 - w_l;         // pygcpic.py:344-346 (mirrored)
    const double Ex = sE[is] * w_l + sE[is + 1] * w_r;
    const double constant = 0.5 * k.dt * cs * 1.602e-19 / m;              // :478
    vx += constant * Ex;
    const double tx = constant * k.B[0], ty = constant * k.B[1], tz = constant * k.B[2];
    const double t2 = tx * tx + ty * ty + tz * tz;
    const double sx = 2. * tx / (1. + t2), sy = 2. * ty / (1. + t2), sz = 2. * tz / (1. + t2);
    const double vfx = vx + vy * tz - vz * ty;
    const double vfy = vy + vz * tx - vx * tz;
    const double vfz = vz + vx * ty - vy * tx;
    vx += vfy * sz - vfz * sy;
    vy += vfz * sx - vfx * sz;
    vz += vfx * sy - vfy * sx;
    vx += constant * Ex;
    o.x = x + vx * k.dt; o.vx = vx; o.vy = vy; o.vz = vz;
    o.ps = max((unsigned)__double2hiint(x) - 1u, (unsigned)__double2hiint(o.x) - 1u);
    const double tf = o.x * idx, ff = floor(tf);
    const unsigned f1 = (unsigned)__double2hiint(tf - ff) - PIC_HI_G;
    const double rf = fma(-ff, k.dx, o.x);
    o.cF = (int)ff;
    o.fr = max(f0, f1);
    const double nr = div_const(pc, k.dx, idx), qr = div_const(cs * PIC_E * pc, k.dx, idx);   // gc_weight_k
    const double wR = rf * idx;
    o.nR = nr * wR; o.nL = nr - o.nR; o.qR = qr * wR; o.qL = qr - o.qR;
}

// exact per-particle routine of the mixed kernel (the body of gc_push_boris_k + gc_weight_k)
__device__ __noinline__ int gm_particle_exact(const GCK& k, const GMix& u, const R7& r, long long i, int act,
                                              const double* sE, int8_t* __restrict__ active, int8_t* __restrict__ at_wall,
                                              int8_t* __restrict__ hit_flag, double* __restrict__ n_acc,
                                              double* __restrict__ rho_acc, int* hits) {
    int bad = 0;
    if (act != 1) { if (hit_flag) hit_flag[i] = 0; return 0; }
    GUni one{u.cs[i], u.m[i], u.p2c[i], u.lean, u.tnow};
    bad = gc_particle_exact(k, one, r, i, act, sE, active, at_wall, hit_flag, nullptr, hits);
    if (active[i] == 1 && n_acc) {
        Cell c = cell_dd_fast(r.r[0][i], k.dx, 1.0 / k.dx);
        if (c.iL < 0 || c.iL > k.ng - 2) { ++bad; c.iL = clampi(c.iL, 0, k.ng - 2); }
        const double nr = div_const(one.p2c, k.dx, 1.0 / k.dx), qr = div_const(one.cs * PIC_E * one.p2c, k.dx, 1.0 / k.dx);
        atomicAdd(&n_acc[c.iL], nr * c.wL); atomicAdd(&n_acc[c.iL + 1], nr * c.wR);
        atomicAdd(&rho_acc[c.iL], qr * c.wL); atomicAdd(&rho_acc[c.iL + 1], qr * c.wR);
    }
    return bad;
}

__device__ __forceinline__ void gmwin_add(double* myw, double* __restrict__ nacc, double* __restrict__ racc, int wb, int c,
                                          const GMixO& o) {
    const unsigned d = (unsigned)(c - wb);
    if (d <= (unsigned)(GM_W - 2)) {
        double* p = myw + d * G_T; p[0] += o.nL; p[G_T] += o.nR;
        p += GM_W * G_T; p[0] += o.qL; p[G_T] += o.qR;
    } else {
        atomicAdd(&nacc[c], o.nL); atomicAdd(&nacc[c + 1], o.nR); atomicAdd(&racc[c], o.qL); atomicAdd(&racc[c + 1], o.qR);
    }
}

template <bool DEP>
__global__ void __launch_bounds__(G_T, 1) gc_push_boris_mix_k(const __grid_constant__ GCK k, const GMix u, int nchunks, const R7 r,
                                                               int8_t* __restrict__ active, int8_t* __restrict__ at_wall,
                                                               int8_t* __restrict__ hit_flag, const double* __restrict__ Egrid,
                                                               double* __restrict__ n_acc, double* __restrict__ rho_acc,
                                                               long long* __restrict__ hit_count, int* __restrict__ range_err) {
    extern __shared__ __align__(128) double sm[];
    __shared__ int s_cnt[2];
    const int ng = k.ng;
    const int NP = (ng + 15) & ~15;
    double* sE = sm;
    double* win = sm + NP;                                   // [2 (n, rho)][GM_W][G_T]
    double* ring = win + (DEP ? 2 * GM_W * G_T : 0);         // [warp][stage][x,vx,vy,vz,cs,m,p2c][64] + flags
    unsigned long long* bars = (unsigned long long*)(ring + (G_T / 32) * GM_NST * GM_SD);
    for (int i = threadIdx.x; i < ng; i += G_T) sE[i] = Egrid[i];
    double* myw = win + threadIdx.x;
    if (DEP) {
#pragma unroll
        for (int n = 0; n < 2 * GM_W; ++n) myw[n * G_T] = 0.0;
    }
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wbase = threadIdx.x & ~31;
    const int NOWIN = -0x40000000;
    const double* wring = ring + warp * (GM_NST * GM_SD);
    const uint32_t ring_s = smem_u32(wring);
    const uint32_t bar_s = smem_u32(bars + warp * GM_NST);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < GM_NST; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const double idx = 1.0 / k.dx;
    const unsigned hi_lim = (unsigned)__double2hiint(k.length) - 1u;
    const int my_chunks = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long woff = (long long)warp * (64 * G_ROWS);
    const long long chunk_step = (long long)gridDim.x * G_CHUNK;
    auto issue = [&](long long base, int st) {
        if (elect_one()) {
            const uint32_t dst = ring_s + st * (GM_SD * 8), bar = bar_s + 8 * st;
            mbar_expect_tx(bar, (uint32_t)(GM_SDP * 8 + 64));
            bulk_g2s(dst, r.r[0] + base, 512, bar);
            bulk_g2s(dst + 512, r.r[3] + base, 512, bar);
            bulk_g2s(dst + 1024, r.r[4] + base, 512, bar);
            bulk_g2s(dst + 1536, r.r[5] + base, 512, bar);
            bulk_g2s(dst + 2048, u.cs + base, 512, bar);
            bulk_g2s(dst + 2560, u.m + base, 512, bar);
            bulk_g2s(dst + 3072, u.p2c + base, 512, bar);
            bulk_g2s(dst + GM_SDP * 8, active + base, 64, bar);
        }
    };
    long long cbase = (long long)blockIdx.x * G_CHUNK + woff;
    if (my_chunks > 0) {
#pragma unroll
        for (int s = 0; s < GM_NST; ++s) issue(cbase + 64 * s, s);
    }
    int stage = 0, bad = 0, hits = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (int c = 0; c < my_chunks; ++c, cbase += chunk_step) {
        const bool more = c + 1 < my_chunks;
        int wb = NOWIN;
        long long ci = cbase + 2 * lane;
#pragma unroll 1
        for (int row = 0; row < G_ROWS; ++row, ci += 64) {
            // y, z and the clock are only accumulated: plain loads now, used at the end of the row
            double2 Y = make_double2(0., 0.), Z = Y, T = Y;
            if (!u.lean) { Y = __ldcs((const double2*)(r.r[1] + ci)); Z = __ldcs((const double2*)(r.r[2] + ci)); T = __ldcs((const double2*)(r.r[6] + ci)); }
            mbar_wait(bar_s + 8 * stage, phase);
            const double* sb = wring + stage * GM_SD + 2 * lane;
            const short fl = *(const short*)((const char*)(wring + stage * GM_SD + GM_SDP) + 2 * lane);
            const int acta = (int)(signed char)(fl & 0xff), actb = (int)(signed char)(fl >> 8);
            const double2 X = *(const double2*)sb, VX = *(const double2*)(sb + 64), VY = *(const double2*)(sb + 128);
            const double2 VZ = *(const double2*)(sb + 192), CS = *(const double2*)(sb + 256), MM = *(const double2*)(sb + 320);
            const double2 PC = *(const double2*)(sb + 384);
            const int st_cur = stage;
            if (++stage == GM_NST) { stage = 0; phase ^= 1u; }
            GMixO a, b;
            gm_fast(k, idx, sE, ng, X.x, VX.x, VY.x, VZ.x, CS.x, MM.x, PC.x, a);
            gm_fast(k, idx, sE, ng, X.y, VX.y, VY.y, VZ.y, CS.y, MM.y, PC.y, b);
            const bool ra = (a.fr > PIC_HI_SPAN) | (a.ps >= hi_lim) | (acta != 1);
            const bool rb = (b.fr > PIC_HI_SPAN) | (b.ps >= hi_lim) | (actb != 1);
            if (DEP && row == 0) {
                int nok = __reduce_add_sync(full, (ra ? 0 : 1) + (rb ? 0 : 1));
                int sum = __reduce_add_sync(full, (ra ? 0 : a.cF) + (rb ? 0 : b.cF));
                wb = nok ? sum / nok - (GM_W - 2) / 2 : NOWIN;
            }
            if (!(ra | rb)) {
                __stcs((double2*)(r.r[0] + ci), make_double2(a.x, b.x));
                __stcs((double2*)(r.r[3] + ci), make_double2(a.vx, b.vx));
                __stcs((double2*)(r.r[4] + ci), make_double2(a.vy, b.vy));
                __stcs((double2*)(r.r[5] + ci), make_double2(a.vz, b.vz));
                if (!u.lean) {
                    __stcs((double2*)(r.r[1] + ci), make_double2(Y.x + a.vy * k.dt, Y.y + b.vy * k.dt));
                    __stcs((double2*)(r.r[2] + ci), make_double2(Z.x + a.vz * k.dt, Z.y + b.vz * k.dt));
                    __stcs((double2*)(r.r[6] + ci), make_double2(T.x + k.dt, T.y + k.dt));
                }
                if (DEP) { gmwin_add(myw, n_acc, rho_acc, wb, a.cF, a); gmwin_add(myw, n_acc, rho_acc, wb, b.cF, b); }
            } else {
                if (ra) bad += gm_particle_exact(k, u, r, ci, acta, sE, active, at_wall, hit_flag, DEP ? n_acc : nullptr, rho_acc, &hits);
                else {
                    r.r[0][ci] = a.x; r.r[3][ci] = a.vx; r.r[4][ci] = a.vy; r.r[5][ci] = a.vz;
                    if (!u.lean) { r.r[1][ci] = Y.x + a.vy * k.dt; r.r[2][ci] = Z.x + a.vz * k.dt; r.r[6][ci] = T.x + k.dt; }
                    if (DEP) gmwin_add(myw, n_acc, rho_acc, wb, a.cF, a);
                }
                if (rb) bad += gm_particle_exact(k, u, r, ci + 1, actb, sE, active, at_wall, hit_flag, DEP ? n_acc : nullptr, rho_acc, &hits);
                else {
                    r.r[0][ci + 1] = b.x; r.r[3][ci + 1] = b.vx; r.r[4][ci + 1] = b.vy; r.r[5][ci + 1] = b.vz;
                    if (!u.lean) { r.r[1][ci + 1] = Y.y + b.vy * k.dt; r.r[2][ci + 1] = Z.y + b.vz * k.dt; r.r[6][ci + 1] = T.y + k.dt; }
                    if (DEP) gmwin_add(myw, n_acc, rho_acc, wb, b.cF, b);
                }
            }
            __syncwarp();
            if (row < G_ROWS - GM_NST) issue(cbase + 64 * (row + GM_NST), st_cur);
            else if (more) issue(cbase + chunk_step + 64 * (row + GM_NST - G_ROWS), st_cur);
        }
        __syncwarp();
        if (DEP && wb != NOWIN) {
            // column sums of the warp's 32 private windows of both quantities -> global accumulators
            double s = 0.0;
            const int n = lane >> 1, half = lane & 1;
            if (n < 2 * GM_W) {
                const double* col = win + n * G_T + wbase + half * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) s += col[(j + n) & 15];
            }
            s += __shfl_xor_sync(full, s, 1);
            if (n < 2 * GM_W && half == 0) {
                const int node = wb + (n < GM_W ? n : n - GM_W);
                if (node >= 0 && node < ng && s != 0.0) atomicAdd((n < GM_W ? n_acc : rho_acc) + node, s);
            }
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 2 * GM_W; ++n2) myw[n2 * G_T] = 0.0;
            __syncwarp();
        }
    }
    // the N % G_CHUNK particles behind the last whole chunk: at most one per thread, exact routine
    for (long long i = (long long)nchunks * G_CHUNK + (long long)blockIdx.x * G_T + threadIdx.x; i < k.N; i += (long long)gridDim.x * G_T)
        bad += gm_particle_exact(k, u, r, i, active[i], sE, active, at_wall, hit_flag, DEP ? n_acc : nullptr, rho_acc, &hits);
    if (bad) atomicAdd(&s_cnt[0], bad);
    if (hits) atomicAdd(&s_cnt[1], hits);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0] && range_err) atomicAdd(range_err, s_cnt[0]);
        if (s_cnt[1] && hit_count) atomicAdd((unsigned long long*)hit_count, (unsigned long long)s_cnt[1]);
    }
}
__global__ void gc_tail_mix_k(GCK k, GMix u, R7 r, long long first, int8_t* __restrict__ active, int8_t* __restrict__ at_wall,
                              int8_t* __restrict__ hit_flag, const double* __restrict__ Egrid, double* __restrict__ n_acc,
                              double* __restrict__ rho_acc, long long* __restrict__ hit_count, int* __restrict__ range_err) {
    int bad = 0, hits = 0;
    for (long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x)
        bad += gm_particle_exact(k, u, r, i, active[i], Egrid, active, at_wall, hit_flag, n_acc, rho_acc, &hits);
    if (hits && hit_count) atomicAdd((unsigned long long*)hit_count, (unsigned long long)hits);
    if (bad && range_err) atomicAdd(range_err, bad);
}

// tail of the uniform fused step (N % G_CHUNK particles): the exact per-particle routine
__global__ void gc_tail_uniform_k(GCK k, GUni u, R7 r, long long first, int8_t* __restrict__ active,
                                  int8_t* __restrict__ at_wall, int8_t* __restrict__ hit_flag,
                                  const double* __restrict__ Egrid, double* __restrict__ n_acc,
                                  long long* __restrict__ hit_count, int* __restrict__ range_err) {
    int bad = 0, hits = 0;
    for (long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x)
        bad += gc_particle_exact(k, u, r, i, active[i], Egrid, active, at_wall, hit_flag, n_acc, &hits);
    if (hits && hit_count) atomicAdd((unsigned long long*)hit_count, (unsigned long long)hits);
    if (bad && range_err) atomicAdd(range_err, bad);
}

// Post-push pass of pic_bca_aps' particle loop (pygcpic.py:1509-1541) for every particle:
//  * Monte-Carlo ionisation attempts (attempt_first_ionization :350-395 for Z==1, charge 0;
//    attempt_nth_ionization :397-458 for Z==5, charge < 3): eligibility flag and the probability
//    density**2 * rate * dx * dt / p2c with the CIC-gathered number density; the uniform draw and
//    the decision stay on the host (legacy stream, index order);
//  * mid-domain exit of wall-born particles (:1530-1541): from_wall and L/2-L/8 < x < L/2+L/8 ->
//    active = 0 (flag returned for the tallies);
//  * the particle's deterministic contribution to the running source-ion count of :1544.
// rate[0..3]: np.interp'd rate coefficients [m^3/s] for (Z,charge) = (1,0), (5,0), (5,1), (5,2).
__global__ void gc_post_push_k(const double* __restrict__ x, const double* __restrict__ p2c, const double* __restrict__ cs,
                               const int32_t* __restrict__ Z, const int8_t* __restrict__ from_wall,
                               int8_t* __restrict__ active, const double* __restrict__ ngrid, int ng, double dx, double dt,
                               double length, double r10, double r50, double r51, double r52, int src_Z,
                               double* __restrict__ prob, int8_t* __restrict__ elig, int8_t* __restrict__ midexit,
                               int8_t* __restrict__ contrib, long long N, int* __restrict__ range_err) {
    int bad = 0;
    const double lo = length / 2 - length / 8, hi = length / 2 + length / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const int act = active[i];
        const int z = Z[i];
        const double c = cs[i], X = x[i];
        const bool el = act == 1 && ((z == 1 && c == 0.0) || (z == 5 && c < 3.0));
        double pr = 0.0;
        if (el) {
            Cell cl = cell_dd(X, dx);
            if (cl.iL < 0 || cl.iL > ng - 2) { ++bad; cl.iL = clampi(cl.iL, 0, ng - 2); }
            const double dens = cl.wL * ngrid[cl.iL] + cl.wR * ngrid[cl.iL + 1];
            const double rate = z == 1 ? r10 : (c == 0.0 ? r50 : (c == 1.0 ? r51 : r52));
            pr = dens * dens * rate * dx * dt / p2c[i];
        }
        const bool mx = from_wall[i] != 0 && act == 1 && (lo < X && X < hi);
        if (mx) active[i] = 0;
        prob[i] = pr;
        elig[i] = el ? 1 : 0;
        midexit[i] = mx ? 1 : 0;
        contrib[i] = (z == src_Z && act == 1 && c > 0.0 && !mx) ? 1 : 0;
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

// n_acc -> (n, rho) of a species-uniform store: n = n_acc, rho = (charge_state*e)*n  (pygcpic.py:880-883)
__global__ void gc_uniform_finish_k(const double* __restrict__ n_acc, double* __restrict__ n, double* __restrict__ rho,
                                    int ng, double cs) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ng; i += gridDim.x * blockDim.x) {
        const double v = n_acc[i];
        n[i] = v;
        rho[i] = cs * PIC_E * v;
    }
}

// CIC deposit of the number density of the listed slots (the particles re-activated after a fused
// push): n_acc[iL] += p2c/dx*wl, n_acc[iL+1] += p2c/dx*wr.
__global__ void gc_deposit_idx_k(const double* __restrict__ x, const int64_t* __restrict__ idx, long long M, double p2c,
                                 double dx, int ng, double* __restrict__ n_acc, int* __restrict__ range_err) {
    int bad = 0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_dd(x[idx[j]], dx);
        if (c.iL < 0 || c.iL > ng - 2) { ++bad; c.iL = clampi(c.iL, 0, ng - 2); }
        const double nr = p2c / dx;
        atomicAdd(&n_acc[c.iL], nr * c.wL);
        atomicAdd(&n_acc[c.iL + 1], nr * c.wR);
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

__global__ void gc_apply_bcs_k(const double* __restrict__ x, int8_t* __restrict__ active, int8_t* __restrict__ at_wall,
                               long long N, double length) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        double X = x[i];
        if (X < 0.0 || X > length) { active[i] = 0; at_wall[i] = 1; }
    }
}

// pygcpic.py:530-550
__global__ void gc_to_gc_k(GCK k, R7 r, const double* __restrict__ cs, const double* __restrict__ m,
                           const int8_t* __restrict__ active) {
    const double B2 = k.B[0] * k.B[0] + k.B[1] * k.B[1] + k.B[2] * k.B[2];
    const double sB = sqrt(B2);
    const double b0 = k.B[0] / sB, b1 = k.B[1] / sB, b2 = k.B[2] / sB;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        if (active[i] != 1) continue;
        double v0 = r.r[3][i], v1 = r.r[4][i], v2 = r.r[5][i];
        double vpar_mag = v0 * b0 + v1 * b1 + v2 * b2;
        double p0 = vpar_mag * b0, p1 = vpar_mag * b1, p2 = vpar_mag * b2;
        double c = cs[i], mm = m[i];
        double wc = fabs(c) * PIC_E * sB / mm;
        double q0 = v0 - p0, q1 = v1 - p1, q2 = v2 - p2;
        double vperp_mag = sqrt(q0 * q0 + q1 * q1 + q2 * q2);
        double h0 = q0 / vperp_mag, h1 = q1 / vperp_mag, h2 = q2 / vperp_mag;
        double mu = 0.5 * mm * (vperp_mag * vperp_mag) / sB;
        double rl_mag = vperp_mag / wc;
        double sgn = (c > 0.0) ? 1.0 : ((c < 0.0) ? -1.0 : 0.0);
        double f = -sgn * PIC_E;                       // the reference multiplies by e here
        double x0 = h1 * b2 - h2 * b1, x1 = h2 * b0 - h0 * b2, x2 = h0 * b1 - h1 * b0;   // cross(vperp_hat,b)
        r.r[0][i] = r.r[0][i] - rl_mag * (f * x0);
        r.r[1][i] = r.r[1][i] - rl_mag * (f * x1);
        r.r[2][i] = r.r[2][i] - rl_mag * (f * x2);
        r.r[3][i] = vpar_mag;
        r.r[4][i] = mu;
    }
}

// pygcpic.py:574-595
__global__ void gc_to_6d_k(GCK k, R7 r, const double* __restrict__ cs, const double* __restrict__ m,
                           const int8_t* __restrict__ active, const double* __restrict__ a0,
                           const double* __restrict__ a1, const double* __restrict__ a2) {
    const double B2 = k.B[0] * k.B[0] + k.B[1] * k.B[1] + k.B[2] * k.B[2];
    const double sB = sqrt(B2);
    const double b0 = k.B[0] / sB, b1 = k.B[1] / sB, b2 = k.B[2] / sB;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        if (active[i] != 1) continue;
        double vpar_mag = r.r[3][i], mu = r.r[4][i], mm = m[i];
        double vperp_mag = sqrt(2.0 * mu * sB / mm);
        double wc = fabs(cs[i]) * PIC_E * sB / mm;
        double rl_mag = vperp_mag / wc;
        double A0 = a0[i], A1 = a1[i], A2 = a2[i];
        double adb = A0 * b0 + A1 * b1 + A2 * b2;
        double p0 = A0 - adb, p1 = A1 - adb, p2 = A2 - adb;      // scalar subtracted (reference quirk)
        double pm = sqrt(p0 * p0 + p1 * p1 + p2 * p2);
        double h0 = p0 / pm, h1 = p1 / pm, h2 = p2 / pm;
        r.r[0][i] = r.r[0][i] + rl_mag * h0;
        r.r[1][i] = r.r[1][i] + rl_mag * h1;
        r.r[2][i] = r.r[2][i] + rl_mag * h2;
        double c0 = b1 * h2 - b2 * h1, c1 = b2 * h0 - b0 * h2, c2 = b0 * h1 - b1 * h0;   // cross(b,bperp_hat)
        r.r[3][i] = vpar_mag * b0 + vperp_mag * c0;
        r.r[4][i] = vpar_mag * b1 + vperp_mag * c1;
        r.r[5][i] = vpar_mag * b2 + vperp_mag * c2;
    }
}

struct Eom { double d0, d1, d2, d3; };
__device__ __forceinline__ Eom eom_gc(double r0, double r1, double r2, double r3, double E0, double E1, double E2,
                                      const double* B, double B2, double sB, double b0, double b1, double b2,
                                      double wc) {
    Eom o;
    double rho = r3 / wc;
    o.d0 = (E1 * B[2] - E2 * B[1]) / B2 + r3 * b0;
    o.d1 = (E2 * B[0] - E0 * B[2]) / B2 + r3 * b1;
    o.d2 = (E0 * B[1] - E1 * B[0]) / B2 + r3 * b2;
    o.d3 = (E0 * r0 + E1 * r1 + E2 * r2) / sB / rho;
    return o;
}
// pygcpic.py:598-645 classic RK4 on (X,Y,Z,vpar); mu, r[5] untouched; t += dt
__global__ void gc_push_rk4_k(GCK k, R7 r, const double* __restrict__ cs, const double* __restrict__ m,
                              const int8_t* __restrict__ active, const double* __restrict__ Egrid,
                              int* __restrict__ range_err) {
    const double B2 = k.B[0] * k.B[0] + k.B[1] * k.B[1] + k.B[2] * k.B[2];
    const double sB = sqrt(B2);
    const double b0 = k.B[0] / sB, b1 = k.B[1] / sB, b2 = k.B[2] / sB;
    int bad = 0;
    const double idx_ = 1.0 / k.dx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        if (active[i] != 1) continue;
        double r0 = r.r[0][i], r1 = r.r[1][i], r2 = r.r[2][i], r3 = r.r[3][i];
        double E0 = Egrid ? ((k.flags & 1) ? Egrid[i] : gather_mirrored(Egrid, r0, k.dx, idx_, k.ng, bad)) : 0.0;
        double E1 = k.Eyz[0], E2 = k.Eyz[1];
        double wc = fabs(cs[i]) * PIC_E * sB / m[i];
        const double dt = k.dt;
        Eom f1 = eom_gc(r0, r1, r2, r3, E0, E1, E2, k.B, B2, sB, b0, b1, b2, wc);
        double k10 = dt * f1.d0, k11 = dt * f1.d1, k12 = dt * f1.d2, k13 = dt * f1.d3;
        Eom f2 = eom_gc(r0 + k10 / 2., r1 + k11 / 2., r2 + k12 / 2., r3 + k13 / 2., E0, E1, E2, k.B, B2, sB, b0, b1, b2, wc);
        double k20 = dt * f2.d0, k21 = dt * f2.d1, k22 = dt * f2.d2, k23 = dt * f2.d3;
        Eom f3 = eom_gc(r0 + k20 / 2., r1 + k21 / 2., r2 + k22 / 2., r3 + k23 / 2., E0, E1, E2, k.B, B2, sB, b0, b1, b2, wc);
        double k30 = dt * f3.d0, k31 = dt * f3.d1, k32 = dt * f3.d2, k33 = dt * f3.d3;
        Eom f4 = eom_gc(r0 + k30, r1 + k31, r2 + k32, r3 + k33, E0, E1, E2, k.B, B2, sB, b0, b1, b2, wc);
        double k40 = dt * f4.d0, k41 = dt * f4.d1, k42 = dt * f4.d2, k43 = dt * f4.d3;
        r.r[0][i] = r0 + (k10 + 2. * k20 + 2. * k30 + k40) / 6.;
        r.r[1][i] = r1 + (k11 + 2. * k21 + 2. * k31 + k41) / 6.;
        r.r[2][i] = r2 + (k12 + 2. * k22 + 2. * k32 + k42) / 6.;
        r.r[3][i] = r3 + (k13 + 2. * k23 + 2. * k33 + k43) / 6.;
        r.r[6][i] = r.r[6][i] + dt;
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

// The same RK4 step for a species-uniform store.  E and B are frozen over the step
// (pygcpic.py:606-612), so the ExB drift terms are computed once per particle, and every division
// whose divisor is uniform (B^2, |B|, the cyclotron frequency, 6) is done with div_const (two
// Markstein corrections from the correctly rounded reciprocal: bit-identical to the IEEE division,
// checked on the device by pic_dev_selftest_div for each constant before the first launch); one
// true division per stage (by the per-particle rho) is left.  5 arrays in/out, 80 B/particle.
struct GRk { double B2, sB, b0, b1, b2, wc, yB2, ysB, ywc, y6, c0; };
template <int MINB>
__global__ void __launch_bounds__(256, MINB) gc_push_rk4_uniform_k(GCK k, GRk u, R7 r, const int8_t* __restrict__ active,
                                                            const double* __restrict__ Egrid,
                                                            int* __restrict__ range_err) {
    int bad = 0;
    const double dt = k.dt;
    const double E1 = k.Eyz[0], E2 = k.Eyz[1];
    const double idx = 1.0 / k.dx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        if (active[i] != 1) continue;
        const double r0 = ld_stream(r.r[0] + i), r1 = ld_stream(r.r[1] + i), r2 = ld_stream(r.r[2] + i);
        const double r3 = ld_stream(r.r[3] + i);
        double E0 = 0.0;
        if (Egrid) {
            if (k.flags & 1) E0 = Egrid[i];
            else {
                Cell c = cell_dd_fast(r0, k.dx, idx);
                if (c.iL < 0 || c.iL > k.ng - 2) { ++bad; c.iL = clampi(c.iL, 0, k.ng - 2); }
                const double w_l = c.wR, w_r = 1.0 - w_l;                   // mirrored, pygcpic.py:344-347
                E0 = Egrid[c.iL] * w_l + Egrid[c.iL + 1] * w_r;
            }
        }
        // drift terms of _eom_GC (pygcpic.py:636-638), identical in the four stages
        const double g1 = div_const(E2 * k.B[0] - E0 * k.B[2], u.B2, u.yB2);
        const double g2 = div_const(E0 * k.B[1] - E1 * k.B[0], u.B2, u.yB2);
#define GC_EOM(a0, a1, a2, a3, d0, d1, d2, d3)                                                    \
        do {                                                                                      \
            const double rho_ = div_const((a3), u.wc, u.ywc);                                     \
            d0 = u.c0 + (a3) * u.b0; d1 = g1 + (a3) * u.b1; d2 = g2 + (a3) * u.b2;                \
            d3 = div_const(E0 * (a0) + E1 * (a1) + E2 * (a2), u.sB, u.ysB) / rho_;               \
        } while (0)
        double f0, f1, f2, f3;
        GC_EOM(r0, r1, r2, r3, f0, f1, f2, f3);
        const double k10 = dt * f0, k11 = dt * f1, k12 = dt * f2, k13 = dt * f3;
        GC_EOM(r0 + k10 / 2., r1 + k11 / 2., r2 + k12 / 2., r3 + k13 / 2., f0, f1, f2, f3);
        const double k20 = dt * f0, k21 = dt * f1, k22 = dt * f2, k23 = dt * f3;
        GC_EOM(r0 + k20 / 2., r1 + k21 / 2., r2 + k22 / 2., r3 + k23 / 2., f0, f1, f2, f3);
        const double k30 = dt * f0, k31 = dt * f1, k32 = dt * f2, k33 = dt * f3;
        GC_EOM(r0 + k30, r1 + k31, r2 + k32, r3 + k33, f0, f1, f2, f3);
        const double k40 = dt * f0, k41 = dt * f1, k42 = dt * f2, k43 = dt * f3;
#undef GC_EOM
        st_stream(r.r[0] + i, r0 + div_const(k10 + 2. * k20 + 2. * k30 + k40, 6., u.y6));
        st_stream(r.r[1] + i, r1 + div_const(k11 + 2. * k21 + 2. * k31 + k41, 6., u.y6));
        st_stream(r.r[2] + i, r2 + div_const(k12 + 2. * k22 + 2. * k32 + k42, 6., u.y6));
        st_stream(r.r[3] + i, r3 + div_const(k13 + 2. * k23 + 2. * k33 + k43, 6., u.y6));
        r.r[6][i] = r.r[6][i] + dt;
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

// Two particles per thread (128-bit loads and stores, two independent RK4 dependency chains in
// flight per thread): the step is bound by fp64 issue and by the latency of its four dependent
// divisions per particle, so instruction-level parallelism is what the grid-stride kernel above lacks.
struct Rk4Out { double x0, x1, x2, x3; };
__device__ __forceinline__ Rk4Out rk4_uniform_one(const GCK& k, const GRk& u, double r0, double r1, double r2, double r3,
                                                  double E0, double E1, double E2, double dt) {
    const double g1 = div_const(E2 * k.B[0] - E0 * k.B[2], u.B2, u.yB2);
    const double g2 = div_const(E0 * k.B[1] - E1 * k.B[0], u.B2, u.yB2);
#define GC_EOM(a0, a1, a2, a3, d0, d1, d2, d3)                                                    \
    do {                                                                                          \
        const double rho_ = div_const((a3), u.wc, u.ywc);                                         \
        d0 = u.c0 + (a3) * u.b0; d1 = g1 + (a3) * u.b1; d2 = g2 + (a3) * u.b2;                    \
        d3 = div_const(E0 * (a0) + E1 * (a1) + E2 * (a2), u.sB, u.ysB) / rho_;                   \
    } while (0)
    double f0, f1, f2, f3;
    GC_EOM(r0, r1, r2, r3, f0, f1, f2, f3);
    const double k10 = dt * f0, k11 = dt * f1, k12 = dt * f2, k13 = dt * f3;
    GC_EOM(r0 + k10 / 2., r1 + k11 / 2., r2 + k12 / 2., r3 + k13 / 2., f0, f1, f2, f3);
    const double k20 = dt * f0, k21 = dt * f1, k22 = dt * f2, k23 = dt * f3;
    GC_EOM(r0 + k20 / 2., r1 + k21 / 2., r2 + k22 / 2., r3 + k23 / 2., f0, f1, f2, f3);
    const double k30 = dt * f0, k31 = dt * f1, k32 = dt * f2, k33 = dt * f3;
    GC_EOM(r0 + k30, r1 + k31, r2 + k32, r3 + k33, f0, f1, f2, f3);
    const double k40 = dt * f0, k41 = dt * f1, k42 = dt * f2, k43 = dt * f3;
#undef GC_EOM
    Rk4Out o;
    o.x0 = r0 + div_const(k10 + 2. * k20 + 2. * k30 + k40, 6., u.y6);
    o.x1 = r1 + div_const(k11 + 2. * k21 + 2. * k31 + k41, 6., u.y6);
    o.x2 = r2 + div_const(k12 + 2. * k22 + 2. * k32 + k42, 6., u.y6);
    o.x3 = r3 + div_const(k13 + 2. * k23 + 2. * k33 + k43, 6., u.y6);
    return o;
}
__device__ __forceinline__ double rk4_gather(const GCK& k, const double* __restrict__ Egrid, double r0, long long i,
                                             double idx, int& bad) {
    if (!Egrid) return 0.0;
    if (k.flags & 1) return Egrid[i];
    Cell c = cell_dd_fast(r0, k.dx, idx);
    if (c.iL < 0 || c.iL > k.ng - 2) { ++bad; c.iL = clampi(c.iL, 0, k.ng - 2); }
    const double w_l = c.wR, w_r = 1.0 - w_l;                   // mirrored, pygcpic.py:344-347
    return Egrid[c.iL] * w_l + Egrid[c.iL + 1] * w_r;
}
template <int MINB>
__global__ void __launch_bounds__(256, MINB) gc_push_rk4_uniform_pair_k(GCK k, GRk u, R7 r, const int8_t* __restrict__ active,
                                                                        const double* __restrict__ Egrid, long long npairs,
                                                                        int* __restrict__ range_err) {
    int bad = 0;
    const double dt = k.dt;
    const double E1 = k.Eyz[0], E2 = k.Eyz[1];
    const double idx = 1.0 / k.dx;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
        const long long i = 2 * p;
        const short fl = *(const short*)(active + i);
        const bool aa = (signed char)(fl & 0xff) == 1, ab = (signed char)(fl >> 8) == 1;
        if (!(aa | ab)) continue;
        const double2 R0 = __ldcs((const double2*)(r.r[0] + i)), R1 = __ldcs((const double2*)(r.r[1] + i));
        const double2 R2 = __ldcs((const double2*)(r.r[2] + i)), R3 = __ldcs((const double2*)(r.r[3] + i));
        const double2 T = __ldcs((const double2*)(r.r[6] + i));
        const double Ea = aa ? rk4_gather(k, Egrid, R0.x, i, idx, bad) : 0.0;
        const double Eb = ab ? rk4_gather(k, Egrid, R0.y, i + 1, idx, bad) : 0.0;
        Rk4Out a = rk4_uniform_one(k, u, R0.x, R1.x, R2.x, R3.x, Ea, E1, E2, dt);
        Rk4Out b = rk4_uniform_one(k, u, R0.y, R1.y, R2.y, R3.y, Eb, E1, E2, dt);
        // an inactive particle of the pair keeps its state
        __stcs((double2*)(r.r[0] + i), make_double2(aa ? a.x0 : R0.x, ab ? b.x0 : R0.y));
        __stcs((double2*)(r.r[1] + i), make_double2(aa ? a.x1 : R1.x, ab ? b.x1 : R1.y));
        __stcs((double2*)(r.r[2] + i), make_double2(aa ? a.x2 : R2.x, ab ? b.x2 : R2.y));
        __stcs((double2*)(r.r[3] + i), make_double2(aa ? a.x3 : R3.x, ab ? b.x3 : R3.y));
        __stcs((double2*)(r.r[6] + i), make_double2(aa ? T.x + dt : T.x, ab ? T.y + dt : T.y));
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

// pygcpic.py:889-904.  state = {n0, p_old, initialised}
__global__ void gc_n0_update_k(const double* __restrict__ phi, const double* __restrict__ n,
                               const double* __restrict__ domain, int ng, double Te, double ve, double added,
                               double dt, double* __restrict__ state) {
    __shared__ double scratch[33];
    double s = 0.0, sn = 0.0;
    for (int i = threadIdx.x; i < ng; i += blockDim.x) {
        sn += n[i];
        if (i + 1 < ng) {
            double e0 = exp(phi[i] / Te / 11600.), e1 = exp(phi[i + 1] / Te / 11600.);
            s += (domain[i + 1] - domain[i]) * (e1 + e0) / 2.0;      // np.trapz(eta, domain)
        }
    }
    s = block_reduce<0>(s, scratch);
    sn = block_reduce<0>(sn, scratch);
    if (threadIdx.x == 0) {
        if (state[2] == 0.0) {
            state[1] = s;
            state[0] = 0.9 * (sn / (double)ng);
            state[2] = 1.0;
        } else {
            double p_new = s;
            double q_new = exp(phi[0] / Te / 11600.) + exp(phi[ng - 1] / Te / 11600.);
            double r_new = 2. * added / dt;
            double fn = sqrt(ve * q_new * dt / p_new);
            state[0] = state[0] * ((1.0 - fn) * state[1] / p_new + fn - fn * fn / 4.) + r_new * dt / p_new;
            state[1] = p_new;
        }
    }
}

// ---------------------------------------------------------------- scans / compaction
// Two int32 values per element are scanned together (exclusive).  SCAN_ITEMS elements
// per thread, 256 threads per CTA.
#define SCAN_T 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_T * SCAN_ITEMS)

struct I2 { int a, b; };
__device__ __forceinline__ I2 block_excl_scan2(I2 v, I2* total, I2* sh /*33*/) {
    unsigned full = 0xffffffffu;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    I2 t = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int ya = __shfl_up_sync(full, t.a, o), yb = __shfl_up_sync(full, t.b, o);
        if (lane >= o) { t.a += ya; t.b += yb; }
    }
    __syncthreads();
    if (lane == 31) sh[w] = t;
    __syncthreads();
    if (w == 0) {
        I2 s = lane < (SCAN_T / 32) ? sh[lane] : I2{0, 0};
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int ya = __shfl_up_sync(full, s.a, o), yb = __shfl_up_sync(full, s.b, o);
            if (lane >= o) { s.a += ya; s.b += yb; }
        }
        sh[lane] = s;
    }
    __syncthreads();
    I2 pre = w > 0 ? sh[w - 1] : I2{0, 0};
    *total = sh[SCAN_T / 32 - 1];
    I2 out;
    out.a = pre.a + t.a - v.a;
    out.b = pre.b + t.b - v.b;
    return out;
}

// element functor: mode 0..2 = compaction predicates on one flag array,
// mode 3 = reactivate-or-delete (a = inactive_entry, b = contrib_after - contrib_entry for actives)
__device__ __forceinline__ I2 scan_value(int mode, const int8_t* f0, const int8_t* f1, const int8_t* f2, long long i) {
    I2 v{0, 0};
    if (mode == 0) v.a = f0[i] != 1;
    else if (mode == 1) v.a = f0[i] == 0;
    else if (mode == 2) v.a = f0[i] != 2;
    else {
        int inact = f0[i] != 0;
        v.a = inact;
        v.b = inact ? 0 : ((int)f2[i] - (int)f1[i]);
    }
    return v;
}

__global__ void __launch_bounds__(SCAN_T) scan_block_sums_k(int mode, const int8_t* __restrict__ f0,
                                                            const int8_t* __restrict__ f1, const int8_t* __restrict__ f2,
                                                            long long N, long long* __restrict__ block_sums) {
    __shared__ I2 sh[33];
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    I2 v{0, 0};
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j)
        if (base + j < N) { I2 e = scan_value(mode, f0, f1, f2, base + j); v.a += e.a; v.b += e.b; }
    I2 total;
    block_excl_scan2(v, &total, sh);
    if (threadIdx.x == 0) { block_sums[2 * (long long)blockIdx.x] = total.a; block_sums[2 * (long long)blockIdx.x + 1] = total.b; }
}
// exclusive scan of the per-CTA sums in place (one CTA, sequential chunks); totals -> block_sums[2*nb..]
__global__ void scan_block_offsets_k(long long* __restrict__ block_sums, int nb) {
    __shared__ long long wa[32], wb[32];
    __shared__ long long ca, cb;
    unsigned full = 0xffffffffu;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) { ca = 0; cb = 0; }
    __syncthreads();
    for (int base = 0; base < nb; base += blockDim.x) {
        int i = base + threadIdx.x;
        long long va = i < nb ? block_sums[2 * i] : 0, vb = i < nb ? block_sums[2 * i + 1] : 0, ta = va, tb = vb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long ya = __shfl_up_sync(full, ta, o), yb = __shfl_up_sync(full, tb, o);
            if (lane >= o) { ta += ya; tb += yb; }
        }
        if (lane == 31) { wa[w] = ta; wb[w] = tb; }
        __syncthreads();
        if (w == 0) {
            long long sa = lane < nw ? wa[lane] : 0, sb = lane < nw ? wb[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long ya = __shfl_up_sync(full, sa, o), yb = __shfl_up_sync(full, sb, o);
                if (lane >= o) { sa += ya; sb += yb; }
            }
            wa[lane] = sa; wb[lane] = sb;
        }
        __syncthreads();
        long long pa = ca + (w > 0 ? wa[w - 1] : 0), pb = cb + (w > 0 ? wb[w - 1] : 0);
        if (i < nb) { block_sums[2 * i] = pa + ta - va; block_sums[2 * i + 1] = pb + tb - vb; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { ca = pa + ta; cb = pb + tb; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { block_sums[2 * (long long)nb] = ca; block_sums[2 * (long long)nb + 1] = cb; }
}
// final pass: write compacted indices (and, for mode 3, the running delta at each inactive slot)
__global__ void __launch_bounds__(SCAN_T) scan_scatter_k(int mode, const int8_t* __restrict__ f0,
                                                         const int8_t* __restrict__ f1, const int8_t* __restrict__ f2,
                                                         long long N, const long long* __restrict__ block_sums, int nb,
                                                         int32_t* __restrict__ idx_out, int32_t* __restrict__ base_out,
                                                         long long* __restrict__ count_out) {
    __shared__ I2 sh[33];
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    I2 e[SCAN_ITEMS];
    I2 v{0, 0};
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        e[j] = (base + j < N) ? scan_value(mode, f0, f1, f2, base + j) : I2{0, 0};
        v.a += e[j].a; v.b += e[j].b;
    }
    I2 total;
    I2 pre = block_excl_scan2(v, &total, sh);
    long long oa = block_sums[2 * (long long)blockIdx.x] + pre.a;
    long long ob = block_sums[2 * (long long)blockIdx.x + 1] + pre.b;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < N && e[j].a) {
            idx_out[oa] = (int32_t)(base + j);
            if (base_out) base_out[oa] = (int32_t)ob;
        }
        oa += e[j].a; ob += e[j].b;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && count_out) count_out[0] = block_sums[2 * (long long)nb];
}

// sequential part of the reactivate-or-delete rule over the K inactive slots only
__global__ void gc_decide_seq_k(const int32_t* __restrict__ idx, const int32_t* __restrict__ base,
                                long long* __restrict__ s /* [0]=K [1]=entry_total -> [2]=react [3]=deleted */,
                                long long source_N, int8_t* __restrict__ decision) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long K = s[0], entry_total = s[1], react = 0, del = 0;
    for (long long t = 0; t < K; ++t) {
        long long cnt = entry_total + (long long)base[t] + react;     // count "at that moment", pygcpic.py:1544
        if (cnt < source_N) { decision[idx[t]] = 1; ++react; }
        else { decision[idx[t]] = 2; ++del; }
    }
    s[2] = react;
    s[3] = del;
}
__global__ void count_flags_k(const int8_t* __restrict__ f, long long N, long long* __restrict__ out) {
    long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) c += f[i] != 0;
    unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(full, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd((unsigned long long*)out, (unsigned long long)c);
}

__global__ void gather_f64_k(const double* __restrict__ src, const int32_t* __restrict__ idx, double* __restrict__ dst,
                             long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        dst[t] = src[idx[t]];
}
// dst[t] = src[perm[t]] for a whole structure of arrays in one pass (the index is read once)
struct SoAPerm {
    const double* sf[12]; double* df[12];
    const int32_t* si[2]; int32_t* di[2];
    const int8_t* sb[6]; int8_t* db[6];
    int nf, ni, nb;
};
__global__ void __launch_bounds__(256) soa_permute_k(const __grid_constant__ SoAPerm a, const int32_t* __restrict__ perm,
                                                     long long n) {
    // four output slots per thread, every load of an array issued before its stores: the gathers are latency
    // bound (a dependent load behind the permutation word), so memory-level parallelism is what counts
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < n; t0 += 4 * stride) {
        long long t[4]; int s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { t[j] = t0 + j * stride; s[j] = t[j] < n ? __ldg(perm + t[j]) : 0; }
        for (int f = 0; f < a.nf; ++f) {
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = t[j] < n ? __ldg(a.sf[f] + s[j]) : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (t[j] < n) __stcs(a.df[f] + t[j], v[j]);
        }
        for (int f = 0; f < a.ni; ++f) {
            int v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = t[j] < n ? __ldg(a.si[f] + s[j]) : 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (t[j] < n) a.di[f][t[j]] = v[j];
        }
        for (int f = 0; f < a.nb; ++f) {
            int8_t v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = t[j] < n ? __ldg(a.sb[f] + s[j]) : (int8_t)0;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (t[j] < n) a.db[f][t[j]] = v[j];
        }
    }
}
__global__ void gather_i8_k(const int8_t* __restrict__ src, const int32_t* __restrict__ idx, int8_t* __restrict__ dst,
                            long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        dst[t] = src[idx[t]];
}

}  // namespace pic

using namespace pic;

static int run_scan(int mode, const int8_t* f0, const int8_t* f1, const int8_t* f2, long long N, int32_t* idx_out,
                    int32_t* base_out, long long* count_out, long long* block_sums, cudaStream_t st) {
    int nb = (int)((N + SCAN_TILE - 1) / SCAN_TILE);
    if (nb < 1) nb = 1;
    scan_block_sums_k<<<nb, SCAN_T, 0, st>>>(mode, f0, f1, f2, N, block_sums);
    PIC_CHECK_LAUNCH();
    scan_block_offsets_k<<<1, 1024, 0, st>>>(block_sums, nb);
    PIC_CHECK_LAUNCH();
    scan_scatter_k<<<nb, SCAN_T, 0, st>>>(mode, f0, f1, f2, N, block_sums, nb, idx_out, base_out, count_out);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

extern "C" {

int pic_dev_gc_interpolate(const double* E, const double* x, double* out, int64_t N, int ng, double dx, int* range_err,
                           void* stream) {
    PIC_REQUIRE(E && x && out && N >= 0 && ng >= 2, "gc_interpolate: bad argument");
    if (N == 0) return PIC_OK;
    gc_interpolate_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(E, x, out, N, ng, dx, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_weight(const double* x, const double* charge_state, const double* p2c, const int8_t* active, double* rho,
                      double* n, int64_t N, int ng, double dx, int* range_err, void* stream) {
    PIC_REQUIRE(x && charge_state && p2c && active && rho && n && N >= 0 && ng >= 2, "gc_weight: bad argument");
    if (N == 0) return PIC_OK;
    size_t smem = (size_t)2 * ng * sizeof(double);
    PIC_REQUIRE(smem <= (size_t)max_optin_smem() - 1024, "gc_weight: ng too large for the shared-memory tiles");
    PIC_CHECK_CUDA(cudaFuncSetAttribute(gc_weight_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gc_weight_k<<<grid_for(N, 256, 4), 256, smem, (cudaStream_t)stream>>>(x, charge_state, p2c, active, rho, n, N, ng, dx,
                                                                          range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_push_boris(const pic_gc_params* p, double* const r[7], const double* charge_state, const double* m,
                          int8_t* active, int8_t* at_wall, int8_t* hit_flag, const double* Egrid, long long* hit_count,
                          int* range_err, void* stream) {
    PIC_REQUIRE(p && r && charge_state && m && active && at_wall && Egrid, "gc_push_boris: null pointer");
    if (p->N == 0) return PIC_OK;
    GCK k = make_gck(p);
    R7 rr;
    for (int i = 0; i < 7; ++i) { PIC_REQUIRE(r[i], "gc_push_boris: null component array"); rr.r[i] = r[i]; }
    size_t smem = (size_t)k.ng * sizeof(double);
    if (!(k.flags & 1) && smem <= (size_t)max_optin_smem() - 1024) k.flags |= 4; else { k.flags &= ~4; smem = 0; }
    PIC_CHECK_CUDA(cudaFuncSetAttribute(gc_push_boris_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem ? smem : 1)));
    gc_push_boris_k<<<grid_for(k.N, 256, 6), 256, smem, (cudaStream_t)stream>>>(k, rr, charge_state, m, active, at_wall,
                                                                                 hit_flag, Egrid, hit_count, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

// Species-uniform fused Boris step (see gc_push_boris_v2_k).  n_acc == NULL: no deposit.
int pic_dev_gc_push_boris_uniform(const pic_gc_params* p, double* const r[7], double charge_state, double m, double p2c,
                                  int8_t* active, int8_t* at_wall, int8_t* hit_flag, const double* Egrid,
                                  double* n_acc, long long* hit_count, int* range_err, void* stream) {
    return pic_dev_gc_push_boris_uniform2(p, r, charge_state, m, p2c, 0, 0.0, active, at_wall, hit_flag, Egrid, n_acc,
                                          hit_count, range_err, stream);
}

int pic_dev_gc_push_boris_uniform2(const pic_gc_params* p, double* const r[7], double charge_state, double m, double p2c,
                                   int lean, double t_now, int8_t* active, int8_t* at_wall, int8_t* hit_flag,
                                   const double* Egrid, double* n_acc, long long* hit_count, int* range_err,
                                   void* stream) {
    PIC_REQUIRE(p && r && active && at_wall && Egrid, "gc_push_boris_uniform: null pointer");
    PIC_REQUIRE(!(p->flags & 3), "gc_push_boris_uniform: pre-gathered E / no-BC modes are not supported");
    PIC_REQUIRE(p->ng >= 8, "gc_push_boris_uniform: grid too small");
    if (p->N == 0) return PIC_OK;
    GCK k = make_gck(p);
    GUni u{charge_state, m, p2c, lean ? 1 : 0, t_now};
    R7 rr;
    bool aligned = true;
    for (int i = 0; i < 7; ++i) {
        PIC_REQUIRE(r[i] || (lean && (i == 1 || i == 2)), "gc_push_boris_uniform: null component array");
        rr.r[i] = r[i];
        aligned = aligned && (((uintptr_t)r[i]) & 15) == 0;
    }
    PIC_REQUIRE(aligned && (((uintptr_t)active) & 15) == 0, "gc_push_boris_uniform: arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const bool dep = n_acc != nullptr;
    auto smem_for = [&](int nst) {
        return ((size_t)((k.ng + 15) & ~15) + (dep ? (size_t)G_W * G_T : 0) + (size_t)(G_T / 32) * nst * ((lean ? 256 : 448) + 16) +
                (size_t)(G_T / 32) * nst) * sizeof(double);
    };
    const size_t cap = (size_t)max_optin_smem() - 512;
    const int nst = lean ? (smem_for(4) <= cap ? 4 : 3) : (smem_for(3) <= cap ? 3 : 2);
    const size_t smem = smem_for(nst);
    PIC_REQUIRE(smem <= cap, "gc_push_boris_uniform: ng too large for the shared-memory field tile");
    const long long nchunks = k.N / G_CHUNK;
    if (nchunks > 0) {
        long long capc = device_sm_count();
        const int grid = (int)(nchunks < capc ? nchunks : capc);
#define PIC_GC_LAUNCH(NST, DEP, ...)                                                                             \
        do {                                                                                                     \
            auto kern = gc_push_boris_v2_k<NST, DEP, ##__VA_ARGS__>;                                             \
            PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            kern<<<grid, G_T, smem, st>>>(k, u, (int)nchunks, rr, active, at_wall, hit_flag, Egrid, n_acc,       \
                                           hit_count, range_err);                                                \
        } while (0)
        if (lean) {
            if (nst == 4) { if (dep) PIC_GC_LAUNCH(4, true, true); else PIC_GC_LAUNCH(4, false, true); }
            else { if (dep) PIC_GC_LAUNCH(3, true, true); else PIC_GC_LAUNCH(3, false, true); }
        } else if (nst == 3) { if (dep) PIC_GC_LAUNCH(3, true); else PIC_GC_LAUNCH(3, false); }
        else { if (dep) PIC_GC_LAUNCH(2, true); else PIC_GC_LAUNCH(2, false); }
#undef PIC_GC_LAUNCH
        PIC_CHECK_LAUNCH();
    }
    const long long done = nchunks * G_CHUNK;
    if (done < k.N && nchunks == 0) gc_tail_uniform_k<<<grid_for(k.N - done, 256, 4), 256, 0, st>>>(k, u, rr, done, active, at_wall, hit_flag, Egrid, n_acc, hit_count, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

// Mixed-store fused Boris step (see gc_push_boris_mix_k).  n_acc == NULL: no deposit (then rho_acc is ignored).
int pic_dev_gc_push_boris_mixed(const pic_gc_params* p, double* const r[7], const double* charge_state, const double* m,
                                const double* p2c, int lean, double t_now, int8_t* active, int8_t* at_wall,
                                int8_t* hit_flag, const double* Egrid, double* n_acc, double* rho_acc,
                                long long* hit_count, int* range_err, void* stream) {
    PIC_REQUIRE(p && r && charge_state && m && p2c && active && at_wall && Egrid, "gc_push_boris_mixed: null pointer");
    PIC_REQUIRE(!(p->flags & 3), "gc_push_boris_mixed: pre-gathered E / no-BC modes are not supported");
    PIC_REQUIRE(p->ng >= 8, "gc_push_boris_mixed: grid too small");
    PIC_REQUIRE(!n_acc || rho_acc, "gc_push_boris_mixed: the fused deposit needs both accumulators");
    if (p->N == 0) return PIC_OK;
    GCK k = make_gck(p);
    GMix u{charge_state, m, p2c, lean ? 1 : 0, t_now};
    R7 rr;
    bool aligned = (((uintptr_t)charge_state | (uintptr_t)m | (uintptr_t)p2c | (uintptr_t)active) & 15) == 0;
    for (int i = 0; i < 7; ++i) {
        PIC_REQUIRE(r[i] || (lean && (i == 1 || i == 2)), "gc_push_boris_mixed: null component array");
        rr.r[i] = r[i];
        aligned = aligned && (((uintptr_t)r[i]) & 15) == 0;
    }
    PIC_REQUIRE(aligned, "gc_push_boris_mixed: arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const bool dep = n_acc != nullptr;
    const size_t smem = ((size_t)((k.ng + 15) & ~15) + (dep ? (size_t)2 * GM_W * G_T : 0) + (size_t)(G_T / 32) * GM_NST * GM_SD +
                         (size_t)(G_T / 32) * GM_NST) * sizeof(double);
    PIC_REQUIRE(smem <= (size_t)max_optin_smem() - 512, "gc_push_boris_mixed: ng too large for the shared-memory field tile");
    const long long nchunks = k.N / G_CHUNK;
    if (nchunks > 0) {
        long long capc = device_sm_count();
        const int grid = (int)(nchunks < capc ? nchunks : capc);
        if (dep) {
            PIC_CHECK_CUDA(cudaFuncSetAttribute(gc_push_boris_mix_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gc_push_boris_mix_k<true><<<grid, G_T, smem, st>>>(k, u, (int)nchunks, rr, active, at_wall, hit_flag, Egrid, n_acc, rho_acc,
                                                               hit_count, range_err);
        } else {
            PIC_CHECK_CUDA(cudaFuncSetAttribute(gc_push_boris_mix_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gc_push_boris_mix_k<false><<<grid, G_T, smem, st>>>(k, u, (int)nchunks, rr, active, at_wall, hit_flag, Egrid, n_acc, rho_acc,
                                                                hit_count, range_err);
        }
        PIC_CHECK_LAUNCH();
    }
    const long long done = nchunks * G_CHUNK;
    if (done < k.N && nchunks == 0) gc_tail_mix_k<<<grid_for(k.N - done, 256, 4), 256, 0, st>>>(k, u, rr, done, active, at_wall, hit_flag, Egrid, n_acc, rho_acc, hit_count, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_post_push(const double* x, const double* p2c, const double* charge_state, const int32_t* Z,
                         const int8_t* from_wall, int8_t* active, const double* n_grid, int ng, double dx, double dt,
                         double length, const double rate[4], int source_Z, double* prob, int8_t* eligible,
                         int8_t* midexit, int8_t* contrib, int64_t N, int* range_err, void* stream) {
    PIC_REQUIRE(N >= 0, "gc_post_push: N<0");
    if (N == 0) return PIC_OK;
    PIC_REQUIRE(x && p2c && charge_state && Z && from_wall && active && n_grid && rate && prob && eligible && midexit &&
                    contrib && ng >= 2,
                "gc_post_push: bad argument");
    gc_post_push_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, p2c, charge_state, Z, from_wall, active, n_grid, ng,
                                                                          dx, dt, length, rate[0], rate[1], rate[2], rate[3],
                                                                          source_Z, prob, eligible, midexit, contrib, N, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_uniform_finish(const double* n_acc, double* n, double* rho, int ng, double charge_state, void* stream) {
    PIC_REQUIRE(n_acc && n && rho && ng >= 2, "gc_uniform_finish: bad argument");
    gc_uniform_finish_k<<<(ng + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_acc, n, rho, ng, charge_state);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_deposit_idx(const double* x, const int64_t* idx, int64_t M, double p2c, double dx, int ng, double* n_acc,
                           int* range_err, void* stream) {
    PIC_REQUIRE(M >= 0, "gc_deposit_idx: M<0");
    if (M == 0) return PIC_OK;
    PIC_REQUIRE(x && idx && n_acc && ng >= 2, "gc_deposit_idx: bad argument");
    gc_deposit_idx_k<<<grid_for(M, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, idx, M, p2c, dx, ng, n_acc, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_apply_bcs(const double* x, int8_t* active, int8_t* at_wall, int64_t N, double length, void* stream) {
    PIC_REQUIRE(x && active && at_wall && N >= 0, "gc_apply_bcs: bad argument");
    if (N == 0) return PIC_OK;
    gc_apply_bcs_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, active, at_wall, N, length);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_to_gc(const pic_gc_params* p, double* const r[7], const double* charge_state, const double* m,
                     const int8_t* active, void* stream) {
    PIC_REQUIRE(p && r && charge_state && m && active, "gc_to_gc: null pointer");
    if (p->N == 0) return PIC_OK;
    GCK k = make_gck(p);
    R7 rr;
    for (int i = 0; i < 7; ++i) rr.r[i] = r[i];
    gc_to_gc_k<<<grid_for(k.N, 256, 8), 256, 0, (cudaStream_t)stream>>>(k, rr, charge_state, m, active);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_to_6d(const pic_gc_params* p, double* const r[7], const double* charge_state, const double* m,
                     const int8_t* active, const double* a0, const double* a1, const double* a2, void* stream) {
    PIC_REQUIRE(p && r && charge_state && m && active && a0 && a1 && a2, "gc_to_6d: null pointer");
    if (p->N == 0) return PIC_OK;
    GCK k = make_gck(p);
    R7 rr;
    for (int i = 0; i < 7; ++i) rr.r[i] = r[i];
    gc_to_6d_k<<<grid_for(k.N, 256, 8), 256, 0, (cudaStream_t)stream>>>(k, rr, charge_state, m, active, a0, a1, a2);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_push_rk4(const pic_gc_params* p, double* const r[7], const double* charge_state, const double* m,
                        const int8_t* active, const double* Egrid, int* range_err, void* stream) {
    PIC_REQUIRE(p && r && charge_state && m && active, "gc_push_rk4: null pointer");
    if (p->N == 0) return PIC_OK;
    GCK k = make_gck(p);
    R7 rr;
    for (int i = 0; i < 7; ++i) rr.r[i] = r[i];
    gc_push_rk4_k<<<grid_for(k.N, 256, 6), 256, 0, (cudaStream_t)stream>>>(k, rr, charge_state, m, active, Egrid, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_push_rk4_uniform(const pic_gc_params* p, double* const r[7], double charge_state, double m,
                                const int8_t* active, const double* Egrid, int* range_err, void* stream) {
    PIC_REQUIRE(p && r && active, "gc_push_rk4_uniform: null pointer");
    if (p->N == 0) return PIC_OK;
    GCK k = make_gck(p);
    R7 rr;
    for (int i = 0; i < 7; ++i) rr.r[i] = r[i];
    GRk u;
    u.B2 = k.B[0] * k.B[0] + k.B[1] * k.B[1] + k.B[2] * k.B[2];
    u.sB = sqrt(u.B2);
    u.b0 = k.B[0] / u.sB; u.b1 = k.B[1] / u.sB; u.b2 = k.B[2] / u.sB;
    u.wc = fabs(charge_state) * PIC_E * u.sB / m;                       // pygcpic.py:629
    u.yB2 = 1.0 / u.B2; u.ysB = 1.0 / u.sB; u.ywc = 1.0 / u.wc; u.y6 = 1.0 / 6.0;
    u.c0 = (k.Eyz[0] * k.B[2] - k.Eyz[1] * k.B[1]) / u.B2;              // :636, uniform
    PIC_REQUIRE(u.B2 > 0.0 && u.wc > 0.0, "gc_push_rk4_uniform: needs B != 0 and a charged species");
    static int minb = -1;
    if (minb < 0) { const char* e = getenv("PIC_RK4_MINB"); minb = e ? atoi(e) : 4; }   // 4 CTAs/SM measured best (2.14 ms per 1e8; 2: 3.43, 3: 2.50, 5: 2.06, 6: 3.01)
    cudaStream_t st = (cudaStream_t)stream;
    static int pair = -1;
    if (pair < 0) { const char* e = getenv("PIC_RK4_PAIR"); pair = e ? atoi(e) : 3; }   // CTAs/SM of the pair kernel; 0: off
    bool al = (((uintptr_t)active) & 1) == 0;
    for (int i = 0; i < 7; ++i) al = al && (((uintptr_t)r[i]) & 15) == 0;
    if (pair > 0 && al && k.N >= 2) {
        const long long npairs = k.N / 2;
        if (pair == 2) gc_push_rk4_uniform_pair_k<2><<<grid_for(npairs, 256, 2), 256, 0, st>>>(k, u, rr, active, Egrid, npairs, range_err);
        else if (pair == 4) gc_push_rk4_uniform_pair_k<4><<<grid_for(npairs, 256, 4), 256, 0, st>>>(k, u, rr, active, Egrid, npairs, range_err);
        else gc_push_rk4_uniform_pair_k<3><<<grid_for(npairs, 256, 3), 256, 0, st>>>(k, u, rr, active, Egrid, npairs, range_err);
        PIC_CHECK_LAUNCH();
        if (k.N % 2 == 0) return PIC_OK;
        // the odd last particle: the one-per-thread kernel on a one-element view
        GCK t = k; t.N = 1;
        R7 r1;
        for (int i = 0; i < 7; ++i) r1.r[i] = rr.r[i] + (k.N - 1);
        gc_push_rk4_uniform_k<4><<<1, 32, 0, st>>>(t, u, r1, active + (k.N - 1), (Egrid && (k.flags & 1)) ? Egrid + (k.N - 1) : Egrid, range_err);
        PIC_CHECK_LAUNCH();
        return PIC_OK;
    }
    if (minb == 2) gc_push_rk4_uniform_k<2><<<grid_for(k.N, 256, 2), 256, 0, st>>>(k, u, rr, active, Egrid, range_err);
    else if (minb == 4) gc_push_rk4_uniform_k<4><<<grid_for(k.N, 256, 4), 256, 0, st>>>(k, u, rr, active, Egrid, range_err);
    else if (minb == 5) gc_push_rk4_uniform_k<5><<<grid_for(k.N, 256, 5), 256, 0, st>>>(k, u, rr, active, Egrid, range_err);
    else if (minb == 6) gc_push_rk4_uniform_k<6><<<grid_for(k.N, 256, 6), 256, 0, st>>>(k, u, rr, active, Egrid, range_err);
    else gc_push_rk4_uniform_k<3><<<grid_for(k.N, 256, 3), 256, 0, st>>>(k, u, rr, active, Egrid, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_n0_update(const double* phi, const double* n, const double* domain, int ng, double Te, double ve,
                         double added_particles, double dt, double* state, void* stream) {
    PIC_REQUIRE(phi && n && domain && state && ng >= 2, "gc_n0_update: bad argument");
    gc_n0_update_k<<<1, 1024, 0, (cudaStream_t)stream>>>(phi, n, domain, ng, Te, ve, added_particles, dt, state);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_decide(const int8_t* inactive_entry, const int8_t* contrib_entry, const int8_t* contrib_after,
                      int8_t* decision, int64_t N, int64_t source_N, int32_t* idx_scratch, int32_t* base_scratch,
                      int64_t* scratch, void* stream) {
    PIC_REQUIRE(inactive_entry && contrib_entry && contrib_after && decision && idx_scratch && base_scratch && scratch,
                "gc_decide: null pointer");
    PIC_REQUIRE(N >= 0 && N < 2147483647LL, "gc_decide: N out of range");
    cudaStream_t st = (cudaStream_t)stream;
    PIC_CHECK_CUDA(cudaMemsetAsync(decision, 0, (size_t)N, st));
    if (N == 0) return PIC_OK;
    // scratch layout: [0] K, [1] entry_total, [2] reactivated, [3] deleted, [4..] per-CTA sums
    long long* s = (long long*)scratch;
    PIC_CHECK_CUDA(cudaMemsetAsync(s, 0, 4 * sizeof(long long), st));
    count_flags_k<<<grid_for(N, 256, 8), 256, 0, st>>>(contrib_entry, N, s + 1);
    PIC_CHECK_LAUNCH();
    int rc = run_scan(3, inactive_entry, contrib_entry, contrib_after, N, idx_scratch, base_scratch, s, s + 4, st);
    if (rc) return rc;
    gc_decide_seq_k<<<1, 32, 0, st>>>(idx_scratch, base_scratch, s, (long long)source_N, decision);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_compact_flags(const int8_t* flags, int64_t N, int mode, int32_t* idx_out, int64_t* count_out,
                          int64_t* block_counts, void* stream) {
    PIC_REQUIRE(flags && idx_out && count_out && block_counts && mode >= 0 && mode <= 2, "compact_flags: bad argument");
    PIC_REQUIRE(N >= 0 && N < 2147483647LL, "compact_flags: N out of range");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) { PIC_CHECK_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int64_t), st)); return PIC_OK; }
    return run_scan(mode, flags, nullptr, nullptr, N, idx_out, nullptr, (long long*)count_out, (long long*)block_counts, st);
}

int pic_dev_gather_f64(const double* src, const int32_t* idx, double* dst, int64_t n, void* stream) {
    PIC_REQUIRE(n >= 0, "gather_f64: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(src && idx && dst, "gather_f64: null pointer");
    gather_f64_k<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}
int pic_dev_soa_permute(const int32_t* perm, int64_t n, const double* const* src_f64, double* const* dst_f64, int n_f64,
                        const int32_t* const* src_i32, int32_t* const* dst_i32, int n_i32, const int8_t* const* src_i8,
                        int8_t* const* dst_i8, int n_i8, void* stream) {
    PIC_REQUIRE(n >= 0 && n_f64 >= 0 && n_f64 <= 12 && n_i32 >= 0 && n_i32 <= 2 && n_i8 >= 0 && n_i8 <= 6,
                "soa_permute: bad array counts");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(perm, "soa_permute: null permutation");
    SoAPerm a;
    memset(&a, 0, sizeof(a));
    a.nf = n_f64; a.ni = n_i32; a.nb = n_i8;
    for (int f = 0; f < n_f64; ++f) { PIC_REQUIRE(src_f64[f] && dst_f64[f], "soa_permute: null f64 array"); a.sf[f] = src_f64[f]; a.df[f] = dst_f64[f]; }
    for (int f = 0; f < n_i32; ++f) { PIC_REQUIRE(src_i32[f] && dst_i32[f], "soa_permute: null i32 array"); a.si[f] = src_i32[f]; a.di[f] = dst_i32[f]; }
    for (int f = 0; f < n_i8; ++f) { PIC_REQUIRE(src_i8[f] && dst_i8[f], "soa_permute: null i8 array"); a.sb[f] = src_i8[f]; a.db[f] = dst_i8[f]; }
    soa_permute_k<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(a, perm, n);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}
int pic_dev_gather_i8(const int8_t* src, const int32_t* idx, int8_t* dst, int64_t n, void* stream) {
    PIC_REQUIRE(n >= 0, "gather_i8: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(src && idx && dst, "gather_i8: null pointer");
    gather_i8_k<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

}  // extern "C"
