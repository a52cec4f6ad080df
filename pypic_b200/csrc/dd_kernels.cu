// PIC_L_DD.py -- bounded two-species implicit (Crank-Nicolson / Picard) sheath.
// The particle phase of one Picard iteration is ONE fused kernel: gather, push,
// wall absorption and the deposition of BOTH currents (jh at the half step, j1 at the
// full step) in a single pass over the structure-of-arrays particle store.
#include <cooperative_groups.h>
#include "common.cuh"
#include "ring.cuh"
#include "host_common.h"
namespace cg = cooperative_groups;

namespace pic {

// slice counters of the dynamically scheduled window kernels: one int per launch, taken round-robin
// from this pool and zeroed on the launch's stream right before it
#define PIC_SCHED_SLOTS 256
__device__ int g_sched_pool[PIC_SCHED_SLOTS];

// optional per-CTA (start,end) %globaltimer pairs for load-balance studies (tools/kbench.py)
__device__ unsigned long long* g_cta_timer = nullptr;
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

struct DDK {
    long long N, n_split;
    int Ng, flags;
    double dx, idx, dt, L, p2c;
    double q[2], c1[2], c2[2];   // c1 = dt*(q/m), c2 = (dt*dt)*(q/m)  (Python evaluation order)
    // reproducible build (flags bit7): every addition to the global accumulators goes to a pair of
    // 64-bit FIXED-POINT words instead (integer addition is associative, so the sums do not depend
    // on which warp, CTA or rank adds first); fix = [hi(2Ng) | lo(2Ng)] behind the fp64 accumulators
    long long* fix;
    int* ferr;                   // device error counter for contributions beyond the fixed-point range
    // enqueue-ahead Picard loop: when non-null and *done != 0 the particle kernels return at once
    // (the field kernel sets it when the residual meets the tolerance), so the host can queue the
    // iterations it expects without a round trip per iteration
    const int* done;
    double fs1, fi1;             // 2^s and 2^-s: one hi unit = 2^-s, one lo unit = 2^-(s+32)
    // absorption log: every particle absorbed by this launch appends {slot, original index, Picard
    // iteration} -- the re-injection then visits the few dead slots instead of scanning N flags, in
    // the reference's index order (PIC_L_DD.py:429-450), and the iteration orders the vionout tally
    // (:497-503).  Absorption is on the rare path only.  dead_buf: int32 [count, 0, 0, 0 | cap x int4]
    int* dead_buf;
    const int* oid;              // original index of the particle in each slot (nullptr: the slot itself)
    int dead_cap, iter;
    long long slot0;             // slot of particle 0 of this launch (tail launches run on offset pointers)
    // FIRST iteration only, optional: mom[0] += sum u0, mom[1] += sum u0*u0 over every particle of the launch --
    // the first iteration streams u0 anyway, so np.std(u0) (PIC_L_DD.py:417) and the kinetic-energy diagnostic
    // (:549) cost no extra pass (the host corrects for the slots re-injected in between, see pic_dev_dd_apply_draws3)
    double* mom;
};
__device__ __forceinline__ void dead_note(const DDK& k, long long i) {
    if (k.dead_buf) {
        const int at = atomicAdd(k.dead_buf, 1);
        const int slot = (int)(k.slot0 + i);
        if (at < k.dead_cap) ((int4*)(k.dead_buf + 4))[at] = make_int4(slot, k.oid ? k.oid[slot] : slot, k.iter, 0);
    }
}

static DDK make_ddk(const pic_dd_params* p) {
    DDK k;
    k.N = p->N; k.n_split = p->n_split; k.Ng = p->Ng; k.flags = p->flags;
    k.dx = p->dx; k.idx = 1. / p->dx; k.dt = p->dt; k.L = p->L; k.p2c = p->p2c;
    for (int s = 0; s < 2; ++s) {
        double qm = p->q[s] / p->m[s];
        k.q[s] = p->q[s];
        k.c1[s] = p->dt * qm;
        k.c2[s] = p->dt * p->dt * qm;
    }
    k.fix = nullptr; k.ferr = nullptr; k.fs1 = 1.0; k.fi1 = 1.0; k.done = nullptr;
    k.dead_buf = nullptr; k.oid = nullptr; k.dead_cap = 0; k.iter = 0; k.slot0 = 0; k.mom = nullptr;
    if (p->flags & 128) {
        // one contribution is q*p2c*u*w/dx with |u| < c: |v| < amax < 2^e, so |v|*2^(31-e) < 2^31 and a
        // node can take 2^31 contributions before the hi word overflows; the lo word carries 32 more bits
        const double qa = fabs(p->q[0]) > fabs(p->q[1]) ? fabs(p->q[0]) : fabs(p->q[1]);
        int e = 0;
        frexp(qa * p->p2c * k.idx * 2.99792458e8, &e);
        k.fs1 = ldexp(1.0, 31 - e); k.fi1 = ldexp(1.0, e - 31);
    }
    return k;
}
// binds the fixed-point words that follow the fp64 accumulators acc[2*Ng+4] (flags bit7)
static void ddk_bind_fix(DDK& k, double* acc, int* range_err) {
    if ((k.flags & 128) && acc) { k.fix = (long long*)(acc + 2 * k.Ng + 4); k.ferr = range_err; }
}

// acc[n] += v on the global accumulators.  Reproducible build: v is split exactly into
// hi = rint(v*2^s) and the remainder, which is rounded to a multiple of 2^-(s+32); both words are
// added with integer atomics, so the result is independent of the order of the additions.
__device__ __forceinline__ void acc_add(const DDK& k, double* __restrict__ acc, int n, double v) {
    if (k.fix) {
        const double t = v * k.fs1, h = rint(t);
        if (!(fabs(h) < 4398046511104.0)) { if (k.ferr) atomicAdd(k.ferr, 1); return; }   // 2^42: 2048 c
        const long long lo = __double2ll_rn((t - h) * 4294967296.0);
        atomicAdd((unsigned long long*)k.fix + n, (unsigned long long)(long long)h);
        atomicAdd((unsigned long long*)k.fix + 2 * k.Ng + n, (unsigned long long)lo);
    } else {
        atomicAdd(&acc[n], v);
    }
}
// fixed-point words of node n -> fp64 (one rounding), words cleared
__device__ __forceinline__ double fix_take(const DDK& k, int n) {
    const long long hi = k.fix[n], lo = k.fix[2 * k.Ng + n];
    k.fix[n] = 0; k.fix[2 * k.Ng + n] = 0;
    return ((double)hi + (double)lo * (1.0 / 4294967296.0)) * k.fi1;
}

// Warp-aggregated deposit of (vL -> node i, vR -> node i+1).  When every lane of the warp
// targets the same cell (the common case once particles are sorted by cell) the warp
// reduces with shuffles and issues ONE pair of atomics; otherwise each lane adds its own.
// GLOB: `tile` is the global accumulator array and off+i the accumulator index (acc_add)
template <bool AGG, bool GLOB = false>
__device__ __forceinline__ void deposit_pair(double* tile, int i, double vL, double vR, bool valid,
                                             const DDK* k = nullptr, int off = 0) {
    if (AGG) {
        unsigned full = 0xffffffffu;
        int key = valid ? i : -1;
        int k0 = __shfl_sync(full, key, 0);
        bool same = __all_sync(full, key == k0);
        if (same) {
            if (k0 < 0) return;
            vL = warp_sum(vL);
            vR = warp_sum(vR);
            if ((threadIdx.x & 31) == 0) {
                if (GLOB) { acc_add(*k, tile, off + i, vL); acc_add(*k, tile, off + i + 1, vR); }
                else { atomicAdd(&tile[i], vL); atomicAdd(&tile[i + 1], vR); }
            }
            return;
        }
    }
    if (valid) {
        if (GLOB) { acc_add(*k, tile, off + i, vL); acc_add(*k, tile, off + i + 1, vR); }
        else { atomicAdd(&tile[i], vL); atomicAdd(&tile[i + 1], vR); }
    }
}

// One Picard iteration, particle phase.  TILE: field + both current tiles live in shared
// memory (3*Ng doubles); otherwise they stay in global memory / L2.
template <bool FIRST, bool TILE, bool AGG>
__global__ void __launch_bounds__(256) dd_picard_iter_k(DDK k, const double* __restrict__ x0,
                                                        const double* __restrict__ u0, const double* x1i, double* x1,
                                                        double* u1, int8_t* __restrict__ active,
                                                        const double* __restrict__ Es, double* __restrict__ acc,
                                                        int* __restrict__ range_err) {
    extern __shared__ double sm[];
    if (k.done && *(const volatile int*)k.done) return;
    const int Ng = k.Ng;
    const double* F = Es;
    double* jh = acc;
    double* j1 = acc + Ng;
    if (TILE) {
        double* sF = sm;
        for (int i = threadIdx.x; i < Ng; i += blockDim.x) { sF[i] = Es[i]; sm[Ng + i] = 0.0; sm[2 * Ng + i] = 0.0; }
        __syncthreads();
        F = sF; jh = sm + Ng; j1 = sm + 2 * Ng;
    }
    int nL0 = 0, nL1 = 0, nR0 = 0, nR1 = 0, bad = 0;
    double ms1 = 0.0, ms2 = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // the trip count is uniform across the warp so the shuffles in deposit_pair are safe
    const long long nIter = (k.N + stride - 1) / stride;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long itn = 0; itn < nIter; ++itn, i += stride) {
        bool alive = i < k.N;
        int sp = 0;
        double X0 = 0., U0 = 0., X1 = 0., U1 = 0., XH = 0., UH = 0.;
        if (alive) {
            sp = i >= k.n_split;
            if (!FIRST) {
                if (active[i] != 1) {
                    // already absorbed in an earlier iteration of this step: the reference
                    // leaves zeros in x1,u1 (PIC_L_DD.py:459-462)
                    x1[i] = 0.0; if (u1) u1[i] = 0.0;
                    alive = false;
                }
            }
        }
        if (alive) {
            X0 = ld_stream(x0 + i);
            U0 = ld_stream(u0 + i);
            if (FIRST) { ms1 += U0; ms2 += U0 * U0; }
            double xs = FIRST ? X0 : (X0 + ld_stream(x1i + i)) * 0.5;   // xs = xh of the previous iteration
            Cell c = cell_dd(xs, k.dx);
            if (c.iL < 0 || c.iL > Ng - 2) { ++bad; c.iL = clampi(c.iL, 0, Ng - 2); c.iR = c.iL + 1; }
            double Ei = c.wL * F[c.iL] + c.wR * F[c.iR];              // PIC_L_DD.py:38
            X1 = X0 + k.dt * U0 + (sp ? k.c2[1] : k.c2[0]) * Ei * 0.5;                  // :479
            U1 = U0 + (sp ? k.c1[1] : k.c1[0]) * Ei;                                   // :481
            XH = (X0 + X1) * 0.5;                                       // :485
            UH = (U0 + U1) * 0.5;                                       // :487
            st_stream(x1 + i, X1);
            if (u1) st_stream(u1 + i, U1);
            // absorption, right wall first (PIC_L_DD.py:495-504)
            if (X0 >= k.L || XH >= k.L || X1 >= k.L) {
                active[i] = 0; alive = false; dead_note(k, i);
                if (sp) ++nR1; else ++nR0;
            } else if (X0 <= 0.0 || XH <= 0.0 || X1 <= 0.0) {
                active[i] = -1; alive = false; dead_note(k, i);
                if (sp) ++nL1; else ++nL0;
            }
        }
        // CIC deposit of the survivors: q*v*p2c*w*idx in the reference's product order (:53-54)
        Cell ch, cf;
        double hL = 0., hR = 0., fL = 0., fR = 0.;
        ch.iL = 0; cf.iL = 0;
        if (alive) {
            ch = cell_dd(XH, k.dx);
            if (ch.iL < 0 || ch.iL > Ng - 2) { ++bad; ch.iL = clampi(ch.iL, 0, Ng - 2); }
            const double qs = sp ? k.q[1] : k.q[0];
            double qv = qs * UH * k.p2c;
            hL = qv * ch.wL * k.idx; hR = qv * ch.wR * k.idx;
            if (u1) {                 // a light iteration (no velocity store) deposits only jh
                cf = cell_dd(X1, k.dx);
                if (cf.iL < 0 || cf.iL > Ng - 2) { ++bad; cf.iL = clampi(cf.iL, 0, Ng - 2); }
                double qf = qs * U1 * k.p2c;
                fL = qf * cf.wL * k.idx; fR = qf * cf.wR * k.idx;
            }
        }
        if (TILE) {
            deposit_pair<AGG>(jh, ch.iL, hL, hR, alive);
            if (u1) deposit_pair<AGG>(j1, cf.iL, fL, fR, alive);
        } else {
            deposit_pair<AGG, true>(acc, ch.iL, hL, hR, alive, &k, 0);
            if (u1) deposit_pair<AGG, true>(acc, cf.iL, fL, fR, alive, &k, Ng);
        }
    }
    if (TILE) {
        __syncthreads();
        for (int n = threadIdx.x; n < 2 * Ng; n += blockDim.x) {
            double v = sm[Ng + n];
            if (v != 0.0) atomicAdd(&acc[n], v);
        }
    }
    // absorbed-this-iteration counters (exact integers carried as fp64)
    unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nL0 += __shfl_xor_sync(full, nL0, o); nL1 += __shfl_xor_sync(full, nL1, o);
        nR0 += __shfl_xor_sync(full, nR0, o); nR1 += __shfl_xor_sync(full, nR1, o);
        bad += __shfl_xor_sync(full, bad, o);
    }
    if (FIRST && k.mom) {
        ms1 = warp_sum(ms1); ms2 = warp_sum(ms2);
        if ((threadIdx.x & 31) == 0) { atomicAdd(k.mom, ms1); atomicAdd(k.mom + 1, ms2); }
    }
    if ((threadIdx.x & 31) == 0) {
        if (nL0) atomicAdd(&acc[2 * Ng + 0], (double)nL0);
        if (nL1) atomicAdd(&acc[2 * Ng + 1], (double)nL1);
        if (nR0) atomicAdd(&acc[2 * Ng + 2], (double)nR0);
        if (nR1) atomicAdd(&acc[2 * Ng + 3], (double)nR1);
        if (bad && range_err) atomicAdd(range_err, bad);
    }
}

// ---------------------------------------------------------------------------------------
// Default variant ("window"): every CTA walks CONTIGUOUS chunks of the particle store and
// every warp a contiguous 32*V4_ROWS-particle slice of the chunk.  Once the store is sorted
// by cell (pic_dev_dd_sort_by_cell, every S steps) the particles of a slice deposit into a
// handful of adjacent nodes, so each thread keeps a PRIVATE window of V4_W nodes per current
// (layout [node][thread] in shared memory: conflict-free plain load-add-store, no atomics and
// no shuffles per particle).  Once per slice the warp sums its 32 private windows column-wise
// and adds the node totals to the global accumulators with fire-and-forget RED.ADD.F64.
//
// The loop body is a branch-free FAST PATH: every rare condition (cell lookup inside the
// rounding guard band, out-of-range remainder/index, wall absorption, particle already
// absorbed, chunk straddling the species boundary) only sets one predicate, and such a
// particle is redone by dd_particle_slow() with the exact IEEE operations.  Range tests on
// doubles are done on the integer pipe by comparing bit patterns (positive doubles order
// like unsigned integers) to keep the half-rate FP64 pipe for arithmetic.  Loads of row
// r+1 are issued before row r is processed (software pipelining).
#define V4_T 1024
#define V4_W 7
#define V4_ROWS 16

struct SlowOut { int code; int bad; };   // code: 0 deposited, 1..4 absorbed (L sp0, L sp1, R sp0, R sp1), 5 skipped

// Full update of ONE particle with the exact (slow) operations; deposits with shared-memory
// atomics into the CTA's fallback tiles tj = [jh | j1].
__device__ __noinline__ SlowOut dd_particle_slow(const DDK& k, long long i, double X0, double U0, double pX1, int act,
                                                 bool first, const double* sF, double* tj, double* x1,
                                                 double* u1, int8_t* __restrict__ active, bool j1 = true,
                                                 bool glob = false) {      // glob: tj is the global accumulator array
    SlowOut o{0, 0};
    const int Ng = k.Ng;
    if (!first && act != 1) { x1[i] = 0.0; if (u1) u1[i] = 0.0; o.code = 5; return o; }   // reference leaves zeros
    const bool sp = i >= k.n_split;
    double xs = first ? X0 : (X0 + pX1) * 0.5;
    Cell c = cell_dd(xs, k.dx);
    if (c.iL < 0 || c.iL > Ng - 2) { ++o.bad; c.iL = clampi(c.iL, 0, Ng - 2); }
    double Ei = c.wL * sF[c.iL] + c.wR * sF[c.iL + 1];
    double X1 = X0 + k.dt * U0 + (sp ? k.c2[1] : k.c2[0]) * Ei * 0.5;
    double U1 = U0 + (sp ? k.c1[1] : k.c1[0]) * Ei;
    double XH = (X0 + X1) * 0.5;
    double UH = (U0 + U1) * 0.5;
    x1[i] = X1; if (u1) u1[i] = U1;
    if (X0 >= k.L || XH >= k.L || X1 >= k.L) { active[i] = 0; dead_note(k, i); o.code = sp ? 4 : 3; return o; }
    if (X0 <= 0.0 || XH <= 0.0 || X1 <= 0.0) { active[i] = -1; dead_note(k, i); o.code = sp ? 2 : 1; return o; }
    Cell a = cell_dd(XH, k.dx);
    if (a.iL < 0 || a.iL > Ng - 2) { ++o.bad; a.iL = clampi(a.iL, 0, Ng - 2); }
    const double qs = sp ? k.q[1] : k.q[0];
    double qv = qs * UH * k.p2c;
    if (glob) { acc_add(k, tj, a.iL, qv * a.wL * k.idx); acc_add(k, tj, a.iL + 1, qv * a.wR * k.idx); }
    else { atomicAdd(&tj[a.iL], qv * a.wL * k.idx); atomicAdd(&tj[a.iL + 1], qv * a.wR * k.idx); }
    if (j1) {
        Cell b = cell_dd(X1, k.dx);
        if (b.iL < 0 || b.iL > Ng - 2) { ++o.bad; b.iL = clampi(b.iL, 0, Ng - 2); }
        double qf = qs * U1 * k.p2c;
        if (glob) { acc_add(k, tj, Ng + b.iL, qf * b.wL * k.idx); acc_add(k, tj, Ng + b.iL + 1, qf * b.wR * k.idx); }
        else { atomicAdd(&tj[Ng + b.iL], qf * b.wL * k.idx); atomicAdd(&tj[Ng + b.iL + 1], qf * b.wR * k.idx); }
    }
    return o;
}

// per-particle fast path: everything except memory traffic and the window update.
// Returns true when the particle must be redone by dd_particle_slow().  All range tests
// look only at the HIGH 32-bit word of the doubles (integer pipe, immediates): they are
// conservative -- a value within 2^-20 (relative) of a limit is sent to the slow path,
// which decides exactly.
struct FastC {
    double dx, idx, dt, c1, c2, qpi;
    double dx2, idxh, c2h, qpih;          // 2*dx, idx/2, c2/2, qpi/2 (dd_fast6: the halvings folded into neighbouring factors)
    unsigned hi_dx, hi_Lm1, ngm2;
};
struct FastO { double X1, U1, hL, hR, fL, fR; int cH, cF; };


template <bool FIRST>
__device__ __forceinline__ bool dd_fast(const FastC& c, const double* __restrict__ sF, int Ng, double X0, double U0,
                                        double pX1, FastO& o) {
    const double xs = FIRST ? X0 : (X0 + pX1) * 0.5;
    const double ts = xs * c.idx, fs = floor(ts);
    bool rare = ((unsigned)__double2hiint(ts - fs) - PIC_HI_G) > PIC_HI_SPAN;
    const double rs = fma(-fs, c.dx, xs);
    rare |= (unsigned)__double2hiint(rs) >= c.hi_dx;
    const int is = (int)fs;
    rare |= (unsigned)is > c.ngm2;
    const double wRs = div_const(rs, c.dx, c.idx), wLs = 1.0 - wRs;
    const int isc = min(max(is, 0), Ng - 2);
    const double Ei = wLs * sF[isc] + wRs * sF[isc + 1];
    o.X1 = X0 + c.dt * U0 + c.c2 * Ei * 0.5;            // PIC_L_DD.py:479
    o.U1 = U0 + c.c1 * Ei;                               // :481
    const double XH = (X0 + o.X1) * 0.5, UH = (U0 + o.U1) * 0.5;
    // absorbed iff X0 or X1 leaves (0,L): XH lies between them (rounding is monotone)
    rare |= ((unsigned)__double2hiint(X0) - 1u) >= c.hi_Lm1;
    rare |= ((unsigned)__double2hiint(o.X1) - 1u) >= c.hi_Lm1;
    // a particle absorbed in an EARLIER iteration of this step left either its out-of-domain
    // position (the iteration it died in) or the reference's 0.0 (later ones) in x1, so the
    // same range test on the previous x1 finds it without streaming the flag array
    if (!FIRST) rare |= ((unsigned)__double2hiint(pX1) - 1u) >= c.hi_Lm1;
    const double th = XH * c.idx, fh = floor(th);
    rare |= ((unsigned)__double2hiint(th - fh) - PIC_HI_G) > PIC_HI_SPAN;
    const double rh = fma(-fh, c.dx, XH);
    rare |= (unsigned)__double2hiint(rh) >= c.hi_dx;
    o.cH = (int)fh;
    rare |= (unsigned)o.cH > c.ngm2;
    const double tf = o.X1 * c.idx, ff = floor(tf);
    rare |= ((unsigned)__double2hiint(tf - ff) - PIC_HI_G) > PIC_HI_SPAN;
    const double rf = fma(-ff, c.dx, o.X1);
    rare |= (unsigned)__double2hiint(rf) >= c.hi_dx;
    o.cF = (int)ff;
    rare |= (unsigned)o.cF > c.ngm2;
    const double ah = c.qpi * UH, af = c.qpi * o.U1;
    o.hR = ah * (rh * c.idx); o.hL = ah - o.hR;
    o.fR = af * (rf * c.idx); o.fL = af - o.fR;
    return rare;
}

#ifndef V5_T
#define V5_T 512
#endif
#define V5_W 7
#define V5_ROWS 16
#define V5_CHUNK (V5_T * 2 * V5_ROWS)

__device__ __forceinline__ void win_add(double* myw, double* tj, int wb, int tile, int Ng, int c, double vL, double vR) {
    const unsigned d = (unsigned)(c - wb);
    if (d <= (unsigned)(V5_W - 2)) {
        double* p = myw + (tile * V5_W + d) * V5_T;
        p[0] += vL; p[V5_T] += vR;
    } else { atomicAdd(&tj[tile * Ng + c], vL); atomicAdd(&tj[tile * Ng + c + 1], vR); }
}

template <bool FIRST>
__global__ void __launch_bounds__(V5_T, 1) dd_picard_iter_v5_k(
    const __grid_constant__ DDK k, long long nchunks, const double* __restrict__ x0, const double* __restrict__ u0,
    const double* x1i, double* x1, double* u1, int8_t* __restrict__ active, const double* __restrict__ Es,
    double* __restrict__ acc, int* __restrict__ range_err) {
    extern __shared__ double sm[];
    __shared__ int s_cnt[8];
    unsigned long long* const tbuf = g_cta_timer;
    if (tbuf && threadIdx.x == 0) tbuf[2 * blockIdx.x] = gtimer();
    const int Ng = k.Ng;
    double* sF = sm;                 // field tile
    double* tj = sm + Ng;            // fallback tiles jh | j1
    double* win = sm + 3 * Ng;       // private windows [2*V5_W][V5_T]
    for (int i = threadIdx.x; i < Ng; i += V5_T) { sF[i] = Es[i]; tj[i] = 0.0; tj[Ng + i] = 0.0; }
    double* myw = win + threadIdx.x;
#pragma unroll
    for (int n = 0; n < 2 * V5_W; ++n) myw[n * V5_T] = 0.0;
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wbase = threadIdx.x & ~31;
    const int NOWIN = -0x40000000;
    FastC fc;
    fc.dx = k.dx; fc.idx = k.idx; fc.dt = k.dt;
    fc.hi_dx = (unsigned)__double2hiint(k.dx);
    fc.hi_Lm1 = (unsigned)__double2hiint(k.L) - 1u;
    fc.ngm2 = (unsigned)(Ng - 2);
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long cstart = ch * V5_CHUNK;
        // warp-contiguous slice of 64*V5_ROWS particles; lane owns the pair (2*lane, 2*lane+1) of each row
        long long i = cstart + (long long)warp * (64 * V5_ROWS) + 2 * lane;
        int wb = NOWIN;
        double2 nX0 = __ldcs((const double2*)(x0 + i)), nU0 = __ldcs((const double2*)(u0 + i));
        double2 nX1 = make_double2(0., 0.);
        if (!FIRST) nX1 = __ldcs((const double2*)(x1i + i));
#pragma unroll 1
        for (int row = 0; row < V5_ROWS; ++row) {
            const long long ci = i;
            const double2 X0 = nX0, U0 = nU0, pX1 = nX1;
            i += 64;
            if (row + 1 < V5_ROWS) {
                nX0 = __ldcs((const double2*)(x0 + i)); nU0 = __ldcs((const double2*)(u0 + i));
                if (!FIRST) nX1 = __ldcs((const double2*)(x1i + i));
            }
            // one species per 64-particle row on the fast path (warp-uniform constants); only the
            // single row that holds the species boundary is redone particle by particle
            const long long rstart = ci - 2 * lane;
            const bool sp = rstart >= k.n_split;
            const bool straddle = !sp && rstart + 64 > k.n_split;
            fc.c1 = sp ? k.c1[1] : k.c1[0]; fc.c2 = sp ? k.c2[1] : k.c2[0];
            fc.qpi = (sp ? k.q[1] : k.q[0]) * k.p2c * k.idx;
            FastO a, b;
            bool ra = dd_fast<FIRST>(fc, sF, Ng, X0.x, U0.x, pX1.x, a);
            bool rb = dd_fast<FIRST>(fc, sF, Ng, X0.y, U0.y, pX1.y, b);
            if (straddle) { ra = true; rb = true; }
            if (row == 0) {
                // window base: centre on the mean deposit cell of the warp's first row
                int nok = __reduce_add_sync(full, (ra ? 0 : 1) + (rb ? 0 : 1));
                int sum = __reduce_add_sync(full, (ra ? 0 : a.cH) + (rb ? 0 : b.cH));
                if (nok) wb = sum / nok - (V5_W - 2) / 2;
            }
            if (!(ra | rb)) {
                __stcs((double2*)(x1 + ci), make_double2(a.X1, b.X1));
                if (u1) __stcs((double2*)(u1 + ci), make_double2(a.U1, b.U1));
            } else {
                if (ra) {
                    SlowOut o = dd_particle_slow(k, ci, X0.x, U0.x, pX1.x, FIRST ? 1 : (int)active[ci], FIRST, sF, tj, x1, u1, active);
                    if (o.code >= 1 && o.code <= 4) atomicAdd(&s_cnt[o.code], 1);
                    if (o.bad) atomicAdd(&s_cnt[0], o.bad);
                } else { x1[ci] = a.X1; if (u1) u1[ci] = a.U1; }
                if (rb) {
                    SlowOut o = dd_particle_slow(k, ci + 1, X0.y, U0.y, pX1.y, FIRST ? 1 : (int)active[ci + 1], FIRST, sF, tj, x1, u1, active);
                    if (o.code >= 1 && o.code <= 4) atomicAdd(&s_cnt[o.code], 1);
                    if (o.bad) atomicAdd(&s_cnt[0], o.bad);
                } else { x1[ci + 1] = b.X1; if (u1) u1[ci + 1] = b.U1; }
            }
            if (!ra) { win_add(myw, tj, wb, 0, Ng, a.cH, a.hL, a.hR); win_add(myw, tj, wb, 1, Ng, a.cF, a.fL, a.fR); }
            if (!rb) { win_add(myw, tj, wb, 0, Ng, b.cH, b.hL, b.hR); win_add(myw, tj, wb, 1, Ng, b.cF, b.fL, b.fR); }
        }
        // column sums of the warp's 32 private windows -> global accumulators
        __syncwarp();
        if (wb != NOWIN) {
            double s = 0.0;
            const int n = lane >> 1, half = lane & 1;
            if (lane < 4 * V5_W) {
                const double* col = win + n * V5_T + wbase + half * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) s += col[(j + n) & 15];
            }
            s += __shfl_xor_sync(full, s, 1);
            if (lane < 4 * V5_W && half == 0) {
                int node = wb + (n < V5_W ? n : n - V5_W);
                if (node >= 0 && node < Ng && s != 0.0) atomicAdd(&acc[(n < V5_W ? 0 : Ng) + node], s);
            }
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 2 * V5_W; ++n2) myw[n2 * V5_T] = 0.0;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < 2 * Ng; n += V5_T) {
        double v = tj[n];
        if (v != 0.0) atomicAdd(&acc[n], v);
    }
    if (threadIdx.x >= 1 && threadIdx.x <= 4 && s_cnt[threadIdx.x])
        atomicAdd(&acc[2 * Ng + threadIdx.x - 1], (double)s_cnt[threadIdx.x]);
    if (threadIdx.x == 0 && s_cnt[0] && range_err) atomicAdd(range_err, s_cnt[0]);
    if (tbuf && threadIdx.x == 0) tbuf[2 * blockIdx.x + 1] = gtimer();
}

// ---------------------------------------------------------------------------------------
// v6: the same fast path and private windows as v5, but the particle rows are staged through
// shared memory by the TMA unit: one elected lane per warp issues 1-D bulk copies
// (cp.async.bulk global -> shared, completion counted on an mbarrier) into a per-warp ring of
// V6_NST stages, so V6_NST rows (x0,u0,x1 of 64 particles each) are in flight per warp without
// holding a single register -- 16 warps x V6_NST x 1.5 KB outstanding per SM, enough to cover
// the HBM latency at full bandwidth with one 512-thread CTA per SM.  The ring runs continuously
// across the CTA's chunks (no bubble at chunk boundaries).  The rare-path / out-of-window
// deposits go straight to the global accumulators (fire-and-forget RED), which frees the
// shared memory the fallback tiles of v5 used for the ring.
#define V6_T 512
// nodes per private deposit window (per tile, per thread) and ring stages.  The window is centred on the slice's
// mean cell after a sort and the electrons then drift ballistically (~0.13 cell per step at 1 sigma): a
// contribution that has left the window is a global RED.  7 nodes lose ~3 sigma after 8 steps.  Two builds of the
// kernel: WIDE (15 nodes, 3 stages: 2-2.5 % faster over a sort interval and good for 12 steps between sorts,
// profiles/r2_window_width_sheath.txt) when the field tile leaves room for it (Ng <= 4352), else 7 nodes / 4 stages
// (Ng <= ~9000, and the large-grid build).
#ifndef V6_W
#define V6_W 7
#endif
#ifndef V6_NST
#define V6_NST 4
#endif
#define V6_W_WIDE 15
#define V6_NST_WIDE 3
#define V6_ROWS 16
#define V6_CHUNK (V6_T * 2 * V6_ROWS)

template <int W>
__device__ __forceinline__ void win_add6(const DDK& k, double* myw, double* acc, int wb, int tile, int Ng, int c, double vL, double vR) {
    const unsigned d = (unsigned)(c - wb);
    if (d <= (unsigned)(W - 2)) {
        double* p = myw + (tile * W + d) * V6_T;
        p[0] += vL; p[V6_T] += vR;
    } else { acc_add(k, acc, tile * Ng + c, vL); acc_add(k, acc, tile * Ng + c + 1, vR); }
}

// Fast path of one particle, v6 flavour.  Instead of a predicate per rare condition it returns
// two unsigned "distances" that the caller reduces with 3-input max instructions:
//   fr : max over the three cell lookups of hi32(frac) - hi32(2^-20); the lookup is exact iff
//        fr <= PIC_HI_SPAN (then floor(x*idx) is the true floor quotient and the remainder
//        x - floor*dx is exact and inside [0,dx), so no separate remainder/index test is needed)
//   ps : max over X0, X1 (and the previous X1) of hi32(X) - 1; all are strictly inside (0,L)
//        -- hence not absorbed, and every cell index inside the grid -- iff ps < hi32(L) - 1.
__device__ __forceinline__ double floor_frac_hi(double xs, double idx) { const double t = xs * idx; return t - floor(t); }

struct FastO6 { double X1, U1, hL, hR, fL, fR; int cH, cF; unsigned fr, ps; bool emiss; };

// BIG (large-grid build): sF is the warp's window of V6_EW field nodes starting at node eb instead
// of the whole field tile; a gather cell outside it sets o.emiss and the particle is redone by the
// exact routine with the field read from global memory.
#define V6_EW 32
// J1 = false: a "light" iteration that deposits only jh -- j1 (the current at n+1, PIC_L_DD.py:513) is
// only used after the Picard loop, so only the LAST iteration needs it (see SheathSim.picard).
template <bool FIRST, bool BIG, bool J1>
__device__ __forceinline__ void dd_fast6(const FastC& c, const double* __restrict__ sF, int Ng, int eb, double X0,
                                         double U0, double pX1, FastO6& o) {
    // Multiplications by 0.5 are exact, so they commute with every rounding: the midpoints xs = (X0+pX1)*0.5,
    // XH = (X0+X1)*0.5, UH = (U0+U1)*0.5 and the factor 0.5 of :479 are never formed -- the doubled quantity is
    // used with idx/2, 2*dx, c2/2, qpi/2 instead (same bits; four multiplications per particle fewer).
    const double s2 = FIRST ? X0 : (X0 + pX1);          // xs (first iteration) or 2*xs
    const double sdx = FIRST ? c.dx : c.dx2, sidx = FIRST ? c.idx : c.idxh;
    const double ts = s2 * sidx, fs = floor(ts);
    const unsigned f0 = (unsigned)__double2hiint(ts - fs) - PIC_HI_G;
    const double rs = fma(-fs, sdx, s2);
    int isc = min(max((int)fs, 0), Ng - 2);                    // keeps the tile read in bounds for rare particles
    o.emiss = false;
    if (BIG) {
        isc -= eb;
        o.emiss = (unsigned)isc > (unsigned)(V6_EW - 2);
        isc = min(max(isc, 0), V6_EW - 2);
    }
    const double wRs = div_const(rs, sdx, sidx), wLs = 1.0 - wRs;
    const double Ei = wLs * sF[isc] + wRs * sF[isc + 1];
    o.X1 = X0 + c.dt * U0 + c.c2h * Ei;                 // PIC_L_DD.py:479
    o.U1 = U0 + c.c1 * Ei;                               // :481
    const double XH2 = X0 + o.X1, UH2 = U0 + o.U1;      // 2*XH, 2*UH
    const unsigned p0 = (unsigned)__double2hiint(X0) - 1u, p1 = (unsigned)__double2hiint(o.X1) - 1u;
    o.ps = FIRST ? max(p0, p1) : __vimax3_u32(p0, p1, (unsigned)__double2hiint(pX1) - 1u);
    const double th = XH2 * c.idxh, fh = floor(th);
    const unsigned f1 = (unsigned)__double2hiint(th - fh) - PIC_HI_G;
    const double rh2 = fma(-fh, c.dx2, XH2);
    o.cH = (int)fh;
    const double ah = c.qpih * UH2;
    o.hR = ah * (rh2 * c.idxh); o.hL = ah - o.hR;
    if (J1) {
        const double tf = o.X1 * c.idx, ff = floor(tf);
        const unsigned f2 = (unsigned)__double2hiint(tf - ff) - PIC_HI_G;
        const double rf = fma(-ff, c.dx, o.X1);
        o.cF = (int)ff;
        o.fr = __vimax3_u32(f0, f1, f2);
        const double af = c.qpi * o.U1;
        o.fR = af * (rf * c.idx); o.fL = af - o.fR;
    } else {
        o.cF = o.cH; o.fr = max(f0, f1); o.fR = 0.0; o.fL = 0.0;
    }
}

// Non-common cases of one particle, cheapest first: (1) absorbed in an earlier iteration of this
// step -> the reference's zeros; (2) the gather lookup was exact and the particle sits strictly
// inside the domain at entry -> the fast-path X1,U1 are the exact values, the walls are tested
// with the reference's comparisons and a survivor deposits into the window (or the global
// accumulators); (3) everything else -> dd_particle_slow (IEEE divisions).
template <bool FIRST, bool J1, int W>
__device__ __forceinline__ void dd_medium(const DDK& k, const FastC& fc, long long i, double X0, double U0, double pX1,
                                          const FastO6& o, int act, bool straddle, bool sp, const double* sF, double* myw,
                                          int wb, double* __restrict__ acc, int* s_cnt, double* x1,
                                          double* u1, int8_t* __restrict__ active) {
    if (!FIRST && act != 1) { x1[i] = 0.0; if (u1) u1[i] = 0.0; return; }
    const unsigned f0 = (unsigned)__double2hiint(floor_frac_hi(FIRST ? X0 : (X0 + pX1) * 0.5, fc.idx)) - PIC_HI_G;
    const unsigned p0 = (unsigned)__double2hiint(X0) - 1u;
    const unsigned pp = FIRST ? 0u : (unsigned)__double2hiint(pX1) - 1u;
    if (!straddle && !o.emiss && f0 <= PIC_HI_SPAN && p0 < fc.hi_Lm1 && pp < fc.hi_Lm1) {
        const double XH = (X0 + o.X1) * 0.5;
        if (X0 >= k.L || XH >= k.L || o.X1 >= k.L) {                    // PIC_L_DD.py:495-499
            x1[i] = o.X1; if (u1) u1[i] = o.U1; active[i] = 0; dead_note(k, i); atomicAdd(&s_cnt[sp ? 4 : 3], 1); return;
        }
        if (X0 <= 0.0 || XH <= 0.0 || o.X1 <= 0.0) {                     // :500-504
            x1[i] = o.X1; if (u1) u1[i] = o.U1; active[i] = -1; dead_note(k, i); atomicAdd(&s_cnt[sp ? 2 : 1], 1); return;
        }
        if (o.fr <= PIC_HI_SPAN && o.ps < fc.hi_Lm1) {
            x1[i] = o.X1; if (u1) u1[i] = o.U1;
            win_add6<W>(k, myw, acc, wb, 0, k.Ng, o.cH, o.hL, o.hR);
            if (J1) win_add6<W>(k, myw, acc, wb, 1, k.Ng, o.cF, o.fL, o.fR);
            return;
        }
    }
    SlowOut so = dd_particle_slow(k, i, X0, U0, pX1, act, FIRST, sF, acc, x1, u1, active, J1, true);
    if (so.code >= 1 && so.code <= 4) atomicAdd(&s_cnt[so.code], 1);
    if (so.bad) atomicAdd(&s_cnt[0], so.bad);
}

template <bool FIRST, bool WU, bool BIG = false, bool WIDE = false>
__global__ void __launch_bounds__(V6_T, 1) dd_picard_iter_v6_k(
    const __grid_constant__ DDK k, int nchunks_fr, const double* __restrict__ x0, const double* __restrict__ u0,
    const double* x1i, double* x1, double* u1, int8_t* __restrict__ active, const double* __restrict__ Es,
    double* __restrict__ acc, int* __restrict__ range_err, int* __restrict__ sched) {
    extern __shared__ __align__(128) double sm[];
    __shared__ int s_cnt[8];
    constexpr int W = WIDE ? V6_W_WIDE : V6_W, NST = WIDE ? V6_NST_WIDE : V6_NST;
    static_assert(W >= 5 && W <= 16, "one lane per (column, half-warp) in the window flush");
    if (k.done && *(const volatile int*)k.done) return;
    unsigned long long* const tbuf = g_cta_timer;
    if (tbuf && threadIdx.x == 0) tbuf[2 * blockIdx.x] = gtimer();
    const int Ng = k.Ng;
    // field tile (whole grid) or, in the large-grid build, one V6_EW-node window per warp
    const int NgP = BIG ? (V6_T / 32) * V6_EW : ((Ng + 15) & ~15);   // keeps what follows 128-byte aligned
    double* sF = sm;
    double* win = sm + NgP;                          // private windows [2*W][V6_T]
    double* ring = win + 2 * W * V6_T;            // [warp][stage][x0|u0|x1][64]
    unsigned long long* bars = (unsigned long long*)(ring + (V6_T / 32) * NST * 192);   // [warp][stage]
    if (!BIG) for (int i = threadIdx.x; i < Ng; i += V6_T) sF[i] = Es[i];
    double* const wE = sm + (threadIdx.x >> 5) * V6_EW;     // BIG: this warp's field window
    const double* const fE = BIG ? wE : sF;                 // what the fast path gathers from
    const double* const gE = BIG ? Es : sF;                 // what the exact routines gather from
    // BIG: few particles per cell -> a 1024-particle slice spans many cells, so the deposit window
    // (and the field window) is re-centred and flushed every V6_FR rows instead of once per slice
    const int FRm = BIG ? (int)((unsigned)nchunks_fr >> 28) : (V6_ROWS - 1);     // rows per window - 1 (a power of two minus one)
    const int nchunks = nchunks_fr & 0x0fffffff;
    int eb = 0;
    double* myw = win + threadIdx.x;
#pragma unroll
    for (int n = 0; n < 2 * W; ++n) myw[n * V6_T] = 0.0;
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wbase = threadIdx.x & ~31;
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
    FastC fc;
    fc.dx = k.dx; fc.idx = k.idx; fc.dt = k.dt;
    fc.dx2 = k.dx * 2.0; fc.idxh = k.idx * 0.5;
    fc.hi_dx = 0; fc.ngm2 = 0;
    fc.hi_Lm1 = (unsigned)__double2hiint(k.L) - 1u;
    // Work units are 1024-particle SLICES, one warp each.  A warp's first slice is static
    // (warp id), every further one is taken from a global counter (sched, zero at launch): a
    // slice that is slow -- e.g. the re-injected, hence unsorted, slots that cluster where the
    // wall-absorbed particles sat -- then delays one warp, not the CTA and not the kernel.  The
    // next slice is acquired when the current one starts, so the TMA ring can run NST rows ahead
    // across the slice boundary.  flags bit6 keeps the static round-robin of whole chunks.
    const bool dynamic = !(k.flags & 64) && sched != nullptr;
    const int nslices = nchunks * (V6_T / 32);
    const int nwarps_total = (int)gridDim.x * (V6_T / 32);
    auto next_slice = [&](int cur) -> int {
        if (!dynamic) return cur + nwarps_total;       // == the same slice position of the CTA's next chunk
        int v = 0;
        if (lane == 0) v = atomicAdd(sched, 1);
        return __shfl_sync(full, v, 0) + nwarps_total;
    };
    const uint32_t row_bytes = FIRST ? 1024u : 1536u;
    // ---- producer: stateless -- the row requested is always the one NST rows ahead of the row
    // being consumed and goes into the stage that row just drained, so every quantity is derived
    // from the consumer's own loop variables (nothing accumulates across iterations) ----
    auto issue = [&](long long base, int st) {
        if (elect_one()) {
            const uint32_t dst = ring_s + st * 1536, bar = bar_s + 8 * st;
            mbar_expect_tx(bar, row_bytes);
            bulk_g2s(dst, x0 + base, 512, bar);
            bulk_g2s(dst + 512, u0 + base, 512, bar);
            if (!FIRST) bulk_g2s(dst + 1024, x1i + base, 512, bar);
        }
    };
    int cur_slice = (int)blockIdx.x * (V6_T / 32) + warp;
    if (cur_slice < nslices) {
#pragma unroll
        for (int s = 0; s < NST; ++s) issue((long long)cur_slice * (64 * V6_ROWS) + 64 * s, s);
    }
    // column sums of the warp's 32 private windows -> global accumulators, windows cleared
    auto flush_windows = [&](int wbase_node) {
        __syncwarp();
        if (wbase_node != NOWIN) {
            // one lane per (column, half of the warp's 32 windows); the jh tile, then -- only an iteration that
            // deposits j1 ever writes it -- the j1 tile
            const int n = lane >> 1, half = lane & 1;
#pragma unroll
            for (int tile = 0; tile < (WU ? 2 : 1); ++tile) {
                double s = 0.0;
                if (n < W) {
                    const double* col = win + (tile * W + n) * V6_T + wbase + half * 16;
#pragma unroll
                    for (int j = 0; j < 16; ++j) s += col[(j + n) & 15];
                }
                s += __shfl_xor_sync(full, s, 1);
                if (n < W && half == 0) {
                    const int node = wbase_node + n;
                    if (node >= 0 && node < Ng && s != 0.0) acc_add(k, acc, tile * Ng + node, s);
                }
            }
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < (WU ? 2 : 1) * W; ++n2) myw[n2 * V6_T] = 0.0;
            __syncwarp();
        }
    };
    int stage = 0;
    uint32_t phase = 0;
    double ms1 = 0.0, ms2 = 0.0;          // FIRST && k.mom: sum u0, sum u0^2 of this thread's particles
#pragma unroll 1
    while (cur_slice < nslices) {
        const int nxt_slice = next_slice(cur_slice);
        const long long cbase = (long long)cur_slice * (64 * V6_ROWS);      // first particle of this slice
        const long long nbase = (long long)nxt_slice * (64 * V6_ROWS);
        // species of the slice; a slice holding the species boundary re-selects per row
        const bool mixed = cbase < k.n_split && cbase + 64 * V6_ROWS > k.n_split;
        const bool sp_slice = cbase >= k.n_split;
        fc.c1 = sp_slice ? k.c1[1] : k.c1[0]; fc.c2 = sp_slice ? k.c2[1] : k.c2[0];
        fc.qpi = (sp_slice ? k.q[1] : k.q[0]) * k.p2c * k.idx;
        fc.c2h = fc.c2 * 0.5; fc.qpih = fc.qpi * 0.5;
        const bool more = nxt_slice < nslices;
        int wb = NOWIN;
        long long ci = cbase + 2 * lane;
#pragma unroll 1
        for (int row = 0; row < V6_ROWS; ++row, ci += 64) {
            if (BIG && row > 0 && (row & FRm) == 0) flush_windows(wb);
            mbar_wait(bar_s + 8 * stage, phase);
            const double* sb = wring + stage * 192 + 2 * lane;
            const double2 X0 = *(const double2*)sb, U0 = *(const double2*)(sb + 64);
            double2 pX1 = make_double2(0., 0.);
            if (!FIRST) pX1 = *(const double2*)(sb + 128);
            if (FIRST && k.mom) { ms1 += U0.x + U0.y; ms2 += U0.x * U0.x + U0.y * U0.y; }
            const int st_cur = stage;
            if (++stage == NST) { stage = 0; phase ^= 1u; }
            bool straddle = false, sp = sp_slice;
            if (mixed) {
                const long long rstart = cbase + 64 * row;
                sp = rstart >= k.n_split;
                straddle = !sp && rstart + 64 > k.n_split;
                fc.c1 = sp ? k.c1[1] : k.c1[0]; fc.c2 = sp ? k.c2[1] : k.c2[0];
                fc.qpi = (sp ? k.q[1] : k.q[0]) * k.p2c * k.idx;
                fc.c2h = fc.c2 * 0.5; fc.qpih = fc.qpi * 0.5;
            }
            if (BIG && (row & FRm) == 0) {
                // field window around the gather cell of the row's first particle
                const double xf = __shfl_sync(full, FIRST ? X0.x : (X0.x + pX1.x) * 0.5, 0);
                const int cb = (int)floor(xf * k.idx);
                eb = min(max(cb - V6_EW / 4, 0), Ng - V6_EW);
                __syncwarp();
                wE[lane] = __ldg(Es + eb + lane);
                __syncwarp();
            }
            FastO6 a, b;
            dd_fast6<FIRST, BIG, WU>(fc, fE, Ng, eb, X0.x, U0.x, pX1.x, a);
            dd_fast6<FIRST, BIG, WU>(fc, fE, Ng, eb, X0.y, U0.y, pX1.y, b);
            const bool ra = (a.fr > PIC_HI_SPAN) | (a.ps >= fc.hi_Lm1) | straddle | a.emiss;
            const bool rb = (b.fr > PIC_HI_SPAN) | (b.ps >= fc.hi_Lm1) | straddle | b.emiss;
            if ((row & FRm) == 0) {
                // window base: centre on the mean deposit cell of the warp's first row
                int nok = __reduce_add_sync(full, (ra ? 0 : 1) + (rb ? 0 : 1));
                int sum = __reduce_add_sync(full, (ra ? 0 : a.cH) + (rb ? 0 : b.cH));
                wb = nok ? sum / nok - (W - 2) / 2 : NOWIN;
            }
            const unsigned dah = (unsigned)(a.cH - wb), daf = (unsigned)(a.cF - wb);
            const unsigned dbh = (unsigned)(b.cH - wb), dbf = (unsigned)(b.cF - wb);
            const bool inwin = (WU ? max(__vimax3_u32(dah, daf, dbh), dbf) : max(dah, dbh)) <= (unsigned)(W - 2);
            if (!(ra | rb)) {
                // the common case: two 128-bit streaming stores and eight conflict-free private RMWs;
                // a lane whose particles drifted out of the warp's window since the last sort falls
                // back to global REDs for those contributions only
                __stcs((double2*)(x1 + ci), make_double2(a.X1, b.X1));
                if (WU) __stcs((double2*)(u1 + ci), make_double2(a.U1, b.U1));
                if (inwin) {
                    double* p = myw + dah * V6_T; p[0] += a.hL; p[V6_T] += a.hR;
                    if (WU) { p = myw + (W + daf) * V6_T; p[0] += a.fL; p[V6_T] += a.fR; }
                    p = myw + dbh * V6_T; p[0] += b.hL; p[V6_T] += b.hR;
                    if (WU) { p = myw + (W + dbf) * V6_T; p[0] += b.fL; p[V6_T] += b.fR; }
                } else {
                    win_add6<W>(k, myw, acc, wb, 0, Ng, a.cH, a.hL, a.hR);
                    if (WU) win_add6<W>(k, myw, acc, wb, 1, Ng, a.cF, a.fL, a.fR);
                    win_add6<W>(k, myw, acc, wb, 0, Ng, b.cH, b.hL, b.hR);
                    if (WU) win_add6<W>(k, myw, acc, wb, 1, Ng, b.cF, b.fL, b.fR);
                }
            } else {
                // one flag load for the pair (ci is even); dead and freshly absorbed particles are
                // settled here without the IEEE-division path
                int acta = 1, actb = 1;
                if (!FIRST) {
                    const short fl = *(const short*)(active + ci);
                    acta = (int)(signed char)(fl & 0xff); actb = (int)(signed char)(fl >> 8);
                }
                dd_medium<FIRST, WU, W>(k, fc, ci, X0.x, U0.x, pX1.x, a, acta, straddle, sp, gE, myw, wb, acc, s_cnt, x1, WU ? u1 : nullptr, active);
                dd_medium<FIRST, WU, W>(k, fc, ci + 1, X0.y, U0.y, pX1.y, b, actb, straddle, sp, gE, myw, wb, acc, s_cnt, x1, WU ? u1 : nullptr, active);
            }
            // Refill the stage this row drained with the row NST ahead (possibly in the next chunk).
            // This must not happen before every lane's LDS of the stage has EXECUTED: an LDS can sit
            // in the load/store queue behind a burst of global atomics for longer than a bulk copy
            // takes, and the copy (async proxy) would then overwrite the slot under it.  Every lane
            // has consumed all three loaded vectors by now (the branch above depends on them), and
            // the barrier orders those uses before the elected lane's copy.
            __syncwarp();
            if (row < V6_ROWS - NST) issue(cbase + 64 * (row + NST), st_cur);
            else if (more) issue(nbase + 64 * (row + NST - V6_ROWS), st_cur);
        }
        flush_windows(wb);
        cur_slice = nxt_slice;
    }
    // the N % V6_CHUNK particles behind the last whole chunk: at most one per thread of the grid, by the exact
    // per-particle routine (deposits straight to the global accumulators) -- a second launch for them cost
    // more than the work (profiles/r2_step_gaps.txt)
    for (long long i = (long long)nchunks * V6_CHUNK + (long long)blockIdx.x * V6_T + threadIdx.x; i < k.N;
         i += (long long)gridDim.x * V6_T) {
        const double X0 = x0[i], U0 = u0[i], pX1 = FIRST ? 0.0 : x1i[i];
        if (FIRST && k.mom) { ms1 += U0; ms2 += U0 * U0; }
        SlowOut so = dd_particle_slow(k, i, X0, U0, pX1, FIRST ? 1 : (int)active[i], FIRST, gE, acc, x1, WU ? u1 : nullptr, active, WU, true);
        if (so.code >= 1 && so.code <= 4) atomicAdd(&s_cnt[so.code], 1);
        if (so.bad) atomicAdd(&s_cnt[0], so.bad);
    }
    if (FIRST && k.mom) {
        ms1 = warp_sum(ms1); ms2 = warp_sum(ms2);
        if (lane == 0) { atomicAdd(k.mom, ms1); atomicAdd(k.mom + 1, ms2); }
    }
    __syncthreads();
    if (threadIdx.x >= 1 && threadIdx.x <= 4 && s_cnt[threadIdx.x])
        atomicAdd(&acc[2 * Ng + threadIdx.x - 1], (double)s_cnt[threadIdx.x]);
    if (threadIdx.x == 0 && s_cnt[0] && range_err) atomicAdd(range_err, s_cnt[0]);
    if (tbuf && threadIdx.x == 0) tbuf[2 * blockIdx.x + 1] = gtimer();
}

// exhaustive-style self test of div_const / cell_dd_fast against the IEEE operations
__global__ void selftest_div_k(double b, unsigned long long n, unsigned long long seed,
                               unsigned long long* __restrict__ mism) {
    const double y = 1.0 / b;
    unsigned long long bad = 0;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)t, (uint32_t)(t >> 32), (uint32_t)seed, (uint32_t)(seed >> 32)};
        philox4x32(c, 0x1234567u, 0x89abcdefu);
        // a in [0,b): the remainder range; plus a few values scaled across many cells
        double a = b * (((double)(((uint64_t)c[0] << 20) ^ (c[1] >> 12))) * (1.0 / 4503599627370496.0));
        if (a >= b) a = nextafter(b, 0.0);
        if (div_const(a, b, y) != a / b) ++bad;
        double x = a * (double)(1 + (c[2] & 0xfffff));          // up to ~1e6 cells
        if ((c[3] & 7) == 0) x = b * (double)(c[2] & 0xfffff);    // node-aligned
        if ((c[3] & 7) == 1) x = nextafter(b * (double)(c[2] & 0xfffff), (c[3] & 8) ? 0.0 : INFINITY);
        Cell p = cell_dd(x, b), f = cell_dd_fast(x, b, y);
        if (p.iL != f.iL || p.wR != f.wR || p.wL != f.wL) ++bad;
    }
    if (bad) atomicAdd(mism, bad);
}

// Field phase, one CTA.  See pic_b200.h for the contract.
__global__ void __launch_bounds__(1024) dd_field_update_k(DDK k, double* __restrict__ acc,
                                                          double* __restrict__ wall_cum,
                                                          const double* __restrict__ E0, double* __restrict__ Es,
                                                          double* __restrict__ E1, double* __restrict__ j1o,
                                                          double* __restrict__ stats, double* __restrict__ Es_prev,
                                                          double* __restrict__ rhist, int* __restrict__ ctl,
                                                          double tol, int maxiter) {
    __shared__ double scratch[33];
    __shared__ double wl[2], wr[2];
    const int Ng = k.Ng;
    if (ctl && *(volatile int*)ctl) return;       // the loop already ended (enqueue-ahead mode)
    if (k.fix)      // reproducible build: the currents arrive as fixed-point words
        for (int i = threadIdx.x; i < 2 * Ng; i += blockDim.x) acc[i] += fix_take(k, i);
    if (threadIdx.x < 4) {
        double v = wall_cum[threadIdx.x] + acc[2 * Ng + threadIdx.x];
        wall_cum[threadIdx.x] = v;
        if (threadIdx.x < 2) wl[threadIdx.x] = v; else wr[threadIdx.x - 2] = v;
    }
    __syncthreads();
    // wall-charge current of every absorbed particle (PIC_L_DD.py:58,62): count * value
    double wallL = wl[0] * (k.dx * k.q[0] * k.p2c / k.dt) + wl[1] * (k.dx * k.q[1] * k.p2c / k.dt);
    double wallR = wr[0] * (-k.dx * k.q[0] * k.p2c / k.dt) + wr[1] * (-k.dx * k.q[1] * k.p2c / k.dt);
    double* jh = acc;
    double* j1 = acc + Ng;
    double sh = 0.0, s1 = 0.0;
    // edge fold j[0]+=j[1]; j[-1]+=j[-2] uses the unfolded neighbours (:65-66)
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        double a = jh[i], b = j1[i];
        if (i == 0) { a = (a + wallL) + jh[1]; b = (b + wallL) + j1[1]; }
        if (i == Ng - 1) { a = (a + wallR) + jh[Ng - 2]; b = (b + wallR) + j1[Ng - 2]; }
        sh += a; s1 += b;
        E1[i] = a;      // staged: folded jh
        j1o[i] = b;
    }
    sh = block_reduce<0>(sh, scratch);
    s1 = block_reduce<0>(s1, scratch);
    const double meanh = sh / (double)Ng;
    const double coef = k.dt / PIC_EPS0;
    double rr = 0.0, ee = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        double e0 = E0[i];
        double e1 = e0 + coef * (meanh - E1[i]);      // :516
        double eh = (e1 + e0) * 0.5;                   // :521
        const double es = Es[i];
        double d = es - eh;
        rr += d * d;
        ee += PIC_EPS0 * e1 * e1 * k.dx / 2.;
        E1[i] = e1;
        if (Es_prev) Es_prev[i] = es;                  // the field this iteration gathered with
        Es[i] = eh;
    }
    rr = block_reduce<0>(rr, scratch);
    ee = block_reduce<0>(ee, scratch);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Ng + 4; i += blockDim.x) acc[i] = 0.0;
    if (threadIdx.x == 0) {
        const double r = sqrt(rr);         // np.linalg.norm(Es-Eh), :525
        const double it = stats[3] + 1.0;
        stats[0] = r;
        stats[1] = s1 / (double)Ng;        // np.average(j1) -> jbias, :551
        stats[2] = ee;                     // sum(eps0*E*E*dx/2), :548
        stats[3] = it;
        if (rhist && it <= (double)maxiter) rhist[(int)it - 1] = r;
        if (ctl && (!(r > tol) || it >= (double)maxiter)) *ctl = 1;      // `while r > tol and k < maxiter`, :452
    }
}

// The same field phase as a device function for the kernel fused with the peer-memory reduction
// below (kept textually separate from dd_field_update_k, the validated default, until the fused
// kernel has been through the same GPU parity suite; to be merged then).
__device__ __forceinline__ void dd_field_update_body(const DDK& k, double* __restrict__ acc,
                                                     double* __restrict__ wall_cum,
                                                     const double* __restrict__ E0, double* __restrict__ Es,
                                                     double* __restrict__ E1, double* __restrict__ j1o,
                                                     double* __restrict__ stats, double* __restrict__ Es_prev,
                                                     double* __restrict__ rhist, int* __restrict__ ctl,
                                                     double tol, int maxiter, double* scratch, double* wl, double* wr) {
    const int Ng = k.Ng;
    if (k.fix)      // reproducible build: the currents arrive as fixed-point words
        for (int i = threadIdx.x; i < 2 * Ng; i += blockDim.x) acc[i] += fix_take(k, i);
    if (threadIdx.x < 4) {
        double v = wall_cum[threadIdx.x] + acc[2 * Ng + threadIdx.x];
        wall_cum[threadIdx.x] = v;
        if (threadIdx.x < 2) wl[threadIdx.x] = v; else wr[threadIdx.x - 2] = v;
    }
    __syncthreads();
    // wall-charge current of every absorbed particle (PIC_L_DD.py:58,62): count * value
    double wallL = wl[0] * (k.dx * k.q[0] * k.p2c / k.dt) + wl[1] * (k.dx * k.q[1] * k.p2c / k.dt);
    double wallR = wr[0] * (-k.dx * k.q[0] * k.p2c / k.dt) + wr[1] * (-k.dx * k.q[1] * k.p2c / k.dt);
    double* jh = acc;
    double* j1 = acc + Ng;
    double sh = 0.0, s1 = 0.0;
    // edge fold j[0]+=j[1]; j[-1]+=j[-2] uses the unfolded neighbours (:65-66)
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        double a = jh[i], b = j1[i];
        if (i == 0) { a = (a + wallL) + jh[1]; b = (b + wallL) + j1[1]; }
        if (i == Ng - 1) { a = (a + wallR) + jh[Ng - 2]; b = (b + wallR) + j1[Ng - 2]; }
        sh += a; s1 += b;
        E1[i] = a;      // staged: folded jh
        j1o[i] = b;
    }
    sh = block_reduce<0>(sh, scratch);
    s1 = block_reduce<0>(s1, scratch);
    const double meanh = sh / (double)Ng;
    const double coef = k.dt / PIC_EPS0;
    double rr = 0.0, ee = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        double e0 = E0[i];
        double e1 = e0 + coef * (meanh - E1[i]);      // :516
        double eh = (e1 + e0) * 0.5;                   // :521
        const double es = Es[i];
        double d = es - eh;
        rr += d * d;
        ee += PIC_EPS0 * e1 * e1 * k.dx / 2.;
        E1[i] = e1;
        if (Es_prev) Es_prev[i] = es;                  // the field this iteration gathered with
        Es[i] = eh;
    }
    rr = block_reduce<0>(rr, scratch);
    ee = block_reduce<0>(ee, scratch);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Ng + 4; i += blockDim.x) acc[i] = 0.0;
    if (threadIdx.x == 0) {
        const double r = sqrt(rr);         // np.linalg.norm(Es-Eh), :525
        const double it = stats[3] + 1.0;
        stats[0] = r;
        stats[1] = s1 / (double)Ng;        // np.average(j1) -> jbias, :551
        stats[2] = ee;                     // sum(eps0*E*E*dx/2), :548
        stats[3] = it;
        if (rhist && it <= (double)maxiter) rhist[(int)it - 1] = r;
        if (ctl && (!(r > tol) || it >= (double)maxiter)) *ctl = 1;      // `while r > tol and k < maxiter`, :452
    }
}


// ---------------------------------------------------------------------------------------
// Particle decomposition WITHOUT a library collective: the field kernel itself exchanges and sums
// the ranks' grid accumulators over NVLink peer memory.  Every rank owns a buffer that all ranks
// have mapped (CUDA IPC):
//     acc  [nacc]                  what this rank's particle kernels add to
//     inbox[2][world][nacc]        inbox[parity][source rank]: the accumulators every rank PUSHED here
//     ready[2][world]   (uint32)   ready[parity][source] = number of the reduction whose data is complete
//     nreal, cnt        (uint32)   local: reductions done so far; CTA arrival counter
// Reduction number n (counted on the device: launches the Picard loop's flag turned into no-ops are
// not counted, identically on every rank), parity b = n & 1, P2P_NB CTAs:
//   1. push: all CTAs copy acc (128-bit loads) into inbox[b][me] of EVERY rank -- posted stores, no
//      round trip over NVLink -- and zero acc; each CTA fences (system scope) and counts in; the last
//      one stores n into ready[b][me] of every rank (release);
//   2. CTA 0 waits until ready[b][*] of its OWN buffer have reached n (local polls), sums the world
//      inbox slots in rank order from LOCAL memory -- the same bits on every rank, so all ranks take
//      the same Picard exit, exactly as with an all-reduce -- runs the field phase and publishes
//      nreal = n.
// No acknowledgement is needed: a rank can push reduction n+2 (the next use of buffer b) only after
// its own reduction n+1 completed, which needed every peer's push of n+1, which every peer issued
// after it had finished reading reduction n.
// (Round 1's protocol PULLED the peers' accumulators with serialized system-scope loads from one CTA
// and acknowledged every read: slower than NCCL.)
// Waits are bounded (a few seconds): on time-out *err is set and the kernel proceeds, so a peer
// that died cannot hang the GPU.
#define P2P_NB 8
struct P2P {
    double* const* peers;      // device array [world] of the ranks' buffers (own entry = local pointer)
    int rank, world, nacc;
    int* err;
};
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double* p2p_inbox(double* buf, int nacc, int world, int par, int src) {
    return buf + (size_t)nacc * (1 + (size_t)par * world + src);
}
__device__ __forceinline__ unsigned* p2p_ready(double* buf, int nacc, int world) {
    return (unsigned*)(buf + (size_t)nacc * (1 + 2 * (size_t)world));
}
__device__ __forceinline__ unsigned* p2p_state(double* buf, int nacc, int world) { return p2p_ready(buf, nacc, world) + 2 * world; }
// bounded wait: 2^g_p2p_spin_log2 polls of >= 100 ns each (default 2^25: seconds, far beyond any
// start-up skew between ranks -- first-call module loads, host-RNG re-injection, a checkpoint write);
// pic_p2p_set_timeout changes it
__device__ int g_p2p_spin_log2 = 25;
__device__ __noinline__ bool p2p_wait(const unsigned* flag, unsigned n) {
    const long long nspin = 1ll << g_p2p_spin_log2;
    for (long long spin = 0; spin < nspin; ++spin) {
        if ((int)(ld_acquire_sys_u32(flag) - n) >= 0) return true;      // wrap-safe "flag >= n"
        __nanosleep(100);
    }
    return false;
}
// step 1 (every CTA of the grid); returns the reduction number
__device__ __forceinline__ unsigned p2p_push(const P2P& P) {
    double* mine = P.peers[P.rank];
    unsigned* st = p2p_state(mine, P.nacc, P.world);
    const unsigned n = *(volatile unsigned*)st + 1u;
    const int par = (int)(n & 1u);
    const int nv = P.nacc >> 1;                                   // nacc is even (2*Ng + 4)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
        const double2 v = ((const double2*)mine)[i];
        ((double2*)mine)[i] = make_double2(0.0, 0.0);
        for (int r = 0; r < P.world; ++r) ((double2*)p2p_inbox(P.peers[r], P.nacc, P.world, par, P.rank))[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(st + 1, 1u);
        if (old == gridDim.x - 1) {
            st[1] = 0u;
            __threadfence_system();
            for (int r = 0; r < P.world; ++r)
                st_release_sys_u32(p2p_ready(P.peers[r], P.nacc, P.world) + par * P.world + P.rank, n);
        }
    }
    return n;
}
// step 2 (CTA 0): sum[nacc] = sum over ranks in rank order
__device__ __forceinline__ void p2p_collect(const P2P& P, unsigned n, double* __restrict__ sum) {
    double* mine = P.peers[P.rank];
    const int par = (int)(n & 1u), t = threadIdx.x;
    if (t < P.world && !p2p_wait(p2p_ready(mine, P.nacc, P.world) + par * P.world + t, n)) atomicExch(P.err, 1);
    __syncthreads();
    const int nv = P.nacc >> 1;
    for (int i = t; i < nv; i += blockDim.x) {
        double2 s = make_double2(0.0, 0.0);
        for (int r = 0; r < P.world; ++r) {
            const double2 v = __ldcg((const double2*)p2p_inbox(mine, P.nacc, P.world, par, r) + i);    // L2: written by the peers
            s.x += v.x; s.y += v.y;
        }
        ((double2*)sum)[i] = s;
    }
    __syncthreads();
}
__device__ __forceinline__ void p2p_publish(const P2P& P, unsigned n) {
    __syncthreads();
    if (threadIdx.x == 0) *(volatile unsigned*)p2p_state(P.peers[P.rank], P.nacc, P.world) = n;
}
__global__ void __launch_bounds__(1024) dd_field_update_p2p_k(DDK k, P2P P, double* __restrict__ acc_sum,
                                                              double* __restrict__ wall_cum,
                                                              const double* __restrict__ E0, double* __restrict__ Es,
                                                              double* __restrict__ E1, double* __restrict__ j1o,
                                                              double* __restrict__ stats, double* __restrict__ Es_prev,
                                                              double* __restrict__ rhist, int* __restrict__ ctl,
                                                              double tol, int maxiter) {
    __shared__ double scratch[33];
    __shared__ double wl[2], wr[2];
    // the same decision in every CTA and on every rank (ctl is only written at the end of an earlier launch):
    // nobody waits for a rank that left
    if (ctl && *(volatile int*)ctl) return;
    const unsigned n = p2p_push(P);
    if (blockIdx.x != 0) return;
    p2p_collect(P, n, acc_sum);
    dd_field_update_body(k, acc_sum, wall_cum, E0, Es, E1, j1o, stats, Es_prev, rhist, ctl, tol, maxiter, scratch, wl, wr);
    p2p_publish(P, n);
}
// the reduction alone (the j1 repair pass): sum[nacc] = sum over ranks, own accumulators zeroed
__global__ void __launch_bounds__(1024) p2p_reduce_k(P2P P, double* __restrict__ sum) {
    const unsigned n = p2p_push(P);
    if (blockIdx.x != 0) return;
    p2p_collect(P, n, sum);
    p2p_publish(P, n);
}

// ---- function-level drop-ins ---------------------------------------------------------
// The same field phase for grids too large for one CTA (Ng > 32768): cooperative launch, one
// CTA per SM, grid-wide barriers between the phases.  The four sums are formed from per-CTA partial
// sums (g_field_part, written without atomics) that every CTA adds IN CTA ORDER after the barrier,
// so the result does not depend on scheduling (the reproducible build stays bit-reproducible at
// any grid size); relative to the single-CTA kernel the sums are re-associated (round-off level
// differences in the residual).
__device__ double g_field_part[4 * 1024];
__global__ void __launch_bounds__(1024) dd_field_update_big_k(DDK k, double* __restrict__ acc,
                                                              double* __restrict__ wall_cum,
                                                              const double* __restrict__ E0, double* __restrict__ Es,
                                                              double* __restrict__ E1, double* __restrict__ j1o,
                                                              double* __restrict__ stats, double* __restrict__ red,
                                                              double* __restrict__ Es_prev, double* __restrict__ rhist,
                                                              int* __restrict__ ctl, double tol, int maxiter) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double scratch[33];
    __shared__ double s_tot[2];
    const int Ng = k.Ng;
    if (ctl && *(volatile int*)ctl) return;       // read by every CTA before the first grid barrier; set after the last
    if (k.fix) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * Ng; i += gridDim.x * blockDim.x) acc[i] += fix_take(k, i);
        grid.sync();
    }
    const double w0 = wall_cum[0] + acc[2 * Ng + 0], w1 = wall_cum[1] + acc[2 * Ng + 1];
    const double w2 = wall_cum[2] + acc[2 * Ng + 2], w3 = wall_cum[3] + acc[2 * Ng + 3];
    const double wallL = w0 * (k.dx * k.q[0] * k.p2c / k.dt) + w1 * (k.dx * k.q[1] * k.p2c / k.dt);
    const double wallR = w2 * (-k.dx * k.q[0] * k.p2c / k.dt) + w3 * (-k.dx * k.q[1] * k.p2c / k.dt);
    double* jh = acc;
    double* j1 = acc + Ng;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    double sh = 0.0, s1 = 0.0;
    for (int i = gtid; i < Ng; i += gsz) {
        double a = jh[i], b = j1[i];
        if (i == 0) { a = (a + wallL) + jh[1]; b = (b + wallL) + j1[1]; }
        if (i == Ng - 1) { a = (a + wallR) + jh[Ng - 2]; b = (b + wallR) + j1[Ng - 2]; }
        sh += a; s1 += b;
        E1[i] = a;
        j1o[i] = b;
    }
    sh = block_reduce<0>(sh, scratch);
    s1 = block_reduce<0>(s1, scratch);
    if (threadIdx.x == 0) { g_field_part[4 * blockIdx.x + 0] = sh; g_field_part[4 * blockIdx.x + 1] = s1; }
    grid.sync();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int c = 0; c < (int)gridDim.x; ++c) { a += g_field_part[4 * c + 0]; b += g_field_part[4 * c + 1]; }
        s_tot[0] = a; s_tot[1] = b;
    }
    __syncthreads();
    const double meanh = s_tot[0] / (double)Ng;
    const double sum_j1 = s_tot[1];
    const double coef = k.dt / PIC_EPS0;
    double rr = 0.0, ee = 0.0;
    for (int i = gtid; i < Ng; i += gsz) {
        double e0 = E0[i];
        double e1 = e0 + coef * (meanh - E1[i]);
        double eh = (e1 + e0) * 0.5;
        const double es = Es[i];
        double d = es - eh;
        rr += d * d;
        ee += PIC_EPS0 * e1 * e1 * k.dx / 2.;
        E1[i] = e1;
        if (Es_prev) Es_prev[i] = es;
        Es[i] = eh;
    }
    rr = block_reduce<0>(rr, scratch);
    ee = block_reduce<0>(ee, scratch);
    if (threadIdx.x == 0) { g_field_part[4 * blockIdx.x + 2] = rr; g_field_part[4 * blockIdx.x + 3] = ee; }
    grid.sync();
    for (int i = gtid; i < 2 * Ng + 4; i += gsz) acc[i] = 0.0;
    if (gtid == 0) {
        double a = 0.0, b = 0.0;
        for (int c = 0; c < (int)gridDim.x; ++c) { a += g_field_part[4 * c + 2]; b += g_field_part[4 * c + 3]; }
        wall_cum[0] = w0; wall_cum[1] = w1; wall_cum[2] = w2; wall_cum[3] = w3;
        const double r = sqrt(a);
        const double it = stats[3] + 1.0;
        stats[0] = r;
        stats[1] = sum_j1 / (double)Ng;
        stats[2] = b;
        stats[3] = it;
        if (rhist && it <= (double)maxiter) rhist[(int)it - 1] = r;
        if (ctl && (!(r > tol) || it >= (double)maxiter)) *ctl = 1;
    }
    (void)red;
}

__global__ void dd_interpolate_k(const double* __restrict__ F, const double* __restrict__ x,
                                 double* __restrict__ out, long long N, int Ng, double dx,
                                 int* __restrict__ range_err) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        Cell c = cell_dd_fast(x[i], dx, 1. / dx);
        if (c.iL < 0 || c.iL > Ng - 2) { if (range_err) atomicAdd(range_err, 1); c.iL = clampi(c.iL, 0, Ng - 2); c.iR = c.iL + 1; }
        out[i] = c.wL * F[c.iL] + c.wR * F[c.iR];
    }
}

// weightCurrents / weightDensities with per-particle q and fp64 active flags.
// acc layout: [0,Ng) CIC, [Ng] left-wall sum, [Ng+1] right-wall sum (raw, folded by dd_weight_fold_k)
template <bool CURRENT>
__global__ void dd_weight_k(const double* __restrict__ x, const double* __restrict__ q,
                            const double* __restrict__ v, const double* __restrict__ active,
                            double* __restrict__ acc, long long N, int Ng, double dx, double dt, double p2c,
                            int* __restrict__ range_err) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < Ng + 2; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    const double idx = 1. / dx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        double a = active[i];
        if (a == 1.0) {
            Cell c = cell_dd(x[i], dx);
            if (c.iL < 0 || c.iL > Ng - 2) { if (range_err) atomicAdd(range_err, 1); c.iL = clampi(c.iL, 0, Ng - 2); }
            double qv = CURRENT ? q[i] * v[i] * p2c : q[i] * p2c;
            atomicAdd(&sm[c.iL], qv * c.wL * idx);
            atomicAdd(&sm[c.iL + 1], qv * c.wR * idx);
        } else if (CURRENT) {
            if (a == -1.0) atomicAdd(&sm[Ng], dx * q[i] * p2c / dt);
            else if (a == 0.0) atomicAdd(&sm[Ng + 1], -dx * q[i] * p2c / dt);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Ng + 2; i += blockDim.x)
        if (sm[i] != 0.0) atomicAdd(&acc[i], sm[i]);
}
__global__ void dd_weight_fold_k(const double* __restrict__ acc, double* __restrict__ out, int Ng, int current) {
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        double a = acc[i];
        if (current) {
            if (i == 0) a = (a + acc[Ng]) + acc[1];
            if (i == Ng - 1) a = (a + acc[Ng + 1]) + acc[Ng - 2];
        }
        out[i] = a;
    }
}

// ---- re-injection -----------------------------------------------------------------------
__global__ void dd_apply_draws_k(const int32_t* __restrict__ idx, const double* __restrict__ xd,
                                 const double* __restrict__ ud, const double* __restrict__ vd,
                                 const double* __restrict__ wd, long long n, double* __restrict__ x0,
                                 double* __restrict__ u0, double* __restrict__ v0, double* __restrict__ w0,
                                 int8_t* __restrict__ active) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        int i = idx[t];
        x0[i] = xd[t]; u0[i] = ud[t];
        if (v0) v0[i] = vd[t];
        if (w0) w0[i] = wd[t];
        active[i] = 1;
    }
}

__device__ __forceinline__ double u52(uint32_t a, uint32_t b) {
    uint64_t v = ((uint64_t)a << 20) ^ (uint64_t)(b >> 12);
    return ((double)v + 0.5) * (1.0 / 4503599627370496.0);
}

// i: slot; o: index the draws are keyed by and where the passive v,w live (the slot itself, or the
// particle's ORIGINAL index when the store tracks it)
__device__ __forceinline__ void dd_reinject_one(const DDK& k, long long i, double* __restrict__ x0,
                                                double* __restrict__ u0, double* __restrict__ v0,
                                                double* __restrict__ w0, int8_t* __restrict__ active, double s0,
                                                double s1, uint64_t seed, uint64_t step, long long goff,
                                                const int32_t* __restrict__ oid = nullptr) {
    const long long o = oid ? (long long)oid[i] : i;
    uint64_t gid = (uint64_t)(goff + o);
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    uint32_t d[4] = {c[0], c[1], c[2], c[3] ^ 0x80000000u};
    uint32_t g[4] = {c[0], c[1], c[2], c[3] ^ 0x40000000u};
    philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32(d, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32(g, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double twopi = 6.283185307179586;
    double ux = u52(c[0], c[1]);
    double r1 = sqrt(-2.0 * log(u52(c[2], c[3]))), t1 = twopi * u52(d[0], d[1]);   // Box-Muller
    double r2 = sqrt(-2.0 * log(u52(d[2], d[3]))), t2 = twopi * u52(g[0], g[1]);
    double sg = (i >= k.n_split) ? s1 : s0;
    x0[i] = ux * k.L;
    u0[i] = sg * r1 * cos(t1);
    if (v0) v0[o] = sg * r1 * sin(t1);
    if (w0) w0[o] = sg * r2 * cos(t2);
    active[i] = 1;
}

// The flag array is scanned 16 flags per 128-bit load; only the (few) dead slots draw.
__global__ void dd_reinject_philox_k(DDK k, double* __restrict__ x0, double* __restrict__ u0,
                                     double* __restrict__ v0, double* __restrict__ w0,
                                     int8_t* __restrict__ active, double s0, double s1, uint64_t seed,
                                     uint64_t step, long long goff, const int32_t* __restrict__ oid,
                                     const int* __restrict__ log_cnt) {
    // the absorption log named every dead slot and the log kernel handled them: nothing to scan
    if (log_cnt && *log_cnt <= k.dead_cap) return;
    const long long nvec = ((uintptr_t)active & 15) == 0 ? k.N / 16 : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const int4 f = __ldg((const int4*)active + v);
        if (f.x == 0x01010101 && f.y == 0x01010101 && f.z == 0x01010101 && f.w == 0x01010101) continue;
        const int w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if ((signed char)(w[j >> 2] >> (8 * (j & 3))) != 1)
                dd_reinject_one(k, v * 16 + j, x0, u0, v0, w0, active, s0, s1, seed, step, goff, oid);
    }
    for (long long i = nvec * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += stride)
        if (active[i] != 1) dd_reinject_one(k, i, x0, u0, v0, w0, active, s0, s1, seed, step, goff, oid);
}

__global__ void sum_sq_k(const double* __restrict__ u, long long N, double scale, double* __restrict__ out) {
    __shared__ double scratch[33];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        double v = u[i];
        s += scale * v * v;
    }
    s = block_reduce<0>(s, scratch);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

// ---- counting sort by (species, cell) ----------------------------------------------------
__device__ __forceinline__ int dd_sort_key(const DDK& k, double x, long long i) {
    int c = (int)floor(x / k.dx);
    c = clampi(c, 0, k.Ng - 1);
    return c + ((i >= k.n_split) ? k.Ng : 0);
}
__global__ void dd_sort_hist_k(DDK k, const double* __restrict__ x0, int32_t* __restrict__ counts) {
    extern __shared__ int sh[];
    const int nk = 2 * k.Ng;
    for (int i = threadIdx.x; i < nk; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    // warp-aggregated: lanes holding the same key elect one leader that adds their count (a store
    // that is already nearly sorted has one or two keys per warp, i.e. 32-way same-address conflicts)
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nIter = (k.N + stride - 1) / stride;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long it = 0; it < nIter; ++it, i += stride) {
        const int key = i < k.N ? dd_sort_key(k, x0[i], i) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&sh[key], __popc(peers));
    }
    __syncthreads();
    for (int i2 = threadIdx.x; i2 < nk; i2 += blockDim.x)
        if (sh[i2]) atomicAdd(&counts[i2], sh[i2]);
}
// exclusive scan of counts (one CTA, sequential over chunks)
__global__ void dd_sort_scan_k(int32_t* __restrict__ counts, int nk) {
    __shared__ int wsum[32];
    __shared__ int carry;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nk; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < nk ? counts[i] : 0, t = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
        if (lane == 31) wsum[w] = t;
        __syncthreads();
        if (w == 0) {
            int s = lane < nw ? wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            wsum[lane] = s;
        }
        __syncthreads();
        int pre = carry + (w > 0 ? wsum[w - 1] : 0);
        if (i < nk) counts[i] = pre + t - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = pre + t;
        __syncthreads();
    }
}
// scatter: each CTA walks chunks of SORT_CHUNK particles; ranks inside a chunk come from
// shared-memory counters, ONE global reservation per (chunk, key present in the chunk) by the
// thread that drew rank 0, and the counters touched are reset by the same threads -- the cost
// per chunk is O(particles), not O(keys) (a nearly sorted store touches a handful of keys).
#define SORT_T 1024
#define SORT_PER 4
template <bool PERM>   // PERM: `us` is an int32 array receiving the source index of every output slot
__global__ void __launch_bounds__(SORT_T) dd_sort_scatter_k(DDK k, const double* __restrict__ x0,
                                                            const double* __restrict__ u0,
                                                            const double* __restrict__ v0,
                                                            const double* __restrict__ w0, double* __restrict__ xs,
                                                            double* __restrict__ us, double* __restrict__ vs,
                                                            double* __restrict__ ws, int32_t* __restrict__ cursor,
                                                            const int32_t* __restrict__ oid, int32_t* __restrict__ oids) {
    extern __shared__ int sh[];   // [0,nk) count of the chunk, [nk,2nk) global base of the chunk's run
    const int nk = 2 * k.Ng;
    int* cnt = sh;
    int* gbase = sh + nk;
    for (int i = threadIdx.x; i < nk; i += SORT_T) cnt[i] = 0;
    __syncthreads();
    const long long chunk = (long long)SORT_T * SORT_PER;
    for (long long base = (long long)blockIdx.x * chunk; base < k.N; base += (long long)gridDim.x * chunk) {
        int key[SORT_PER], rank[SORT_PER];
        double X[SORT_PER];
#pragma unroll
        for (int j = 0; j < SORT_PER; ++j) {
            const long long i = base + (long long)j * SORT_T + threadIdx.x;
            key[j] = -1; rank[j] = 0; X[j] = 0.;
            if (i < k.N) { X[j] = x0[i]; key[j] = dd_sort_key(k, X[j], i); }
            // warp-aggregated rank: one shared-memory atomic per distinct key in the warp
            const unsigned peers = __match_any_sync(0xffffffffu, key[j]);
            const unsigned lane = threadIdx.x & 31;
            const int leader = __ffs(peers) - 1;
            int rbase = 0;
            if (key[j] >= 0 && (int)lane == leader) rbase = atomicAdd(&cnt[key[j]], __popc(peers));
            rbase = __shfl_sync(0xffffffffu, rbase, leader);
            rank[j] = rbase + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SORT_PER; ++j)
            if (key[j] >= 0 && rank[j] == 0) gbase[key[j]] = atomicAdd(&cursor[key[j]], cnt[key[j]]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SORT_PER; ++j) {
            const long long i = base + (long long)j * SORT_T + threadIdx.x;
            if (key[j] >= 0) {
                const long long pos = (long long)gbase[key[j]] + rank[j];
                xs[pos] = X[j];
                if (PERM) ((int32_t*)us)[pos] = (int32_t)i; else us[pos] = u0[i];
                if (vs) vs[pos] = v0[i];
                if (ws) ws[pos] = w0[i];
                if (oids) oids[pos] = oid ? oid[i] : (int32_t)i;     // original-index payload (identity on the first sort)
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SORT_PER; ++j)
            if (key[j] >= 0 && rank[j] == 0) cnt[key[j]] = 0;
        __syncthreads();
    }
}

// ---- the same counting sort for grids whose histogram does not fit shared memory: counts and
// cursors live in global memory (L2), one warp-aggregated atomic per distinct key per warp ----
// Every CTA walks a CONTIGUOUS range of the store (on a nearly sorted store the CTAs then work on
// different cells, so their atomics go to different addresses instead of all hammering the few
// counters of the cells a grid-stride sweep is crossing), four elements per thread in flight.
#define SORTB_T 256
#define SORTB_PER 4
__device__ __forceinline__ long long sortb_range(long long N, long long* end) {
    long long per = (N + gridDim.x - 1) / gridDim.x;
    per = (per + SORTB_T * SORTB_PER - 1) / (SORTB_T * SORTB_PER) * (SORTB_T * SORTB_PER);
    const long long beg = (long long)blockIdx.x * per;
    *end = beg + per < N ? beg + per : N;
    return beg;
}
__global__ void __launch_bounds__(SORTB_T) dd_sort_hist_big_k(DDK k, const double* __restrict__ x0, int32_t* __restrict__ counts) {
    long long end;
    const long long beg = sortb_range(k.N, &end);
    for (long long base = beg; base < end; base += SORTB_T * SORTB_PER) {
        double X[SORTB_PER];
#pragma unroll
        for (int j = 0; j < SORTB_PER; ++j) {
            const long long i = base + j * SORTB_T + threadIdx.x;
            X[j] = i < end ? x0[i] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < SORTB_PER; ++j) {
            const long long i = base + j * SORTB_T + threadIdx.x;
            const int key = i < end ? dd_sort_key(k, X[j], i) : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (key >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&counts[key], __popc(peers));
        }
    }
}
// exclusive scan in three passes: per-1024-block local scan + block sums, scan of the sums (dd_sort_scan_k), add
__global__ void __launch_bounds__(1024) scan_local_k(int32_t* __restrict__ v, int n, int32_t* __restrict__ sums) {
    __shared__ int wsum[32];
    const int i = blockIdx.x * 1024 + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int x = i < n ? v[i] : 0;
    int t = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
    if (lane == 31) wsum[w] = t;
    __syncthreads();
    if (w == 0) {
        int s2 = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s2, o); if (lane >= o) s2 += y; }
        wsum[lane] = s2;
    }
    __syncthreads();
    const int pre = w > 0 ? wsum[w - 1] : 0;
    if (i < n) v[i] = pre + t - x;
    if (threadIdx.x == 1023) sums[blockIdx.x] = pre + t;
}
__global__ void scan_add_k(int32_t* __restrict__ v, int n, const int32_t* __restrict__ sums) {
    const int i = blockIdx.x * 1024 + threadIdx.x;
    if (i < n) v[i] += sums[blockIdx.x];
}
// ---- STABLE sort by cell (reproducible build): least-significant-digit radix sort of one species
// block, 8-bit digits of the cell index, three launches per pass (per-tile digit histogram laid out
// [digit][tile], exclusive scan of that table, scatter).  The rank of an element inside its tile is
// its position in index order among the tile's elements with the same digit (warp ballot ranks +
// a per-digit prefix over the tile's warps), so equal cells keep their previous order and the
// result does not depend on scheduling -- unlike the counting sort above, whose order inside a cell
// follows the arrival order of atomic reservations.
#define RS_T 1024
#define RS_PER 4
#define RS_TILE (RS_T * RS_PER)
__device__ __forceinline__ int rs_digit(double x, double dx, int Ng, int shift) {
    int c = (int)floor(x / dx);
    c = clampi(c, 0, Ng - 1);
    return (c >> shift) & 255;
}
__global__ void __launch_bounds__(RS_T) rsort_hist_k(const double* __restrict__ x, long long n, double dx, int Ng,
                                                     int shift, int ntiles, int32_t* __restrict__ H) {
    __shared__ int hist[256];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        const long long base = (long long)tile * RS_TILE;
#pragma unroll
        for (int j = 0; j < RS_PER; ++j) {
            const long long i = base + (long long)j * RS_T + threadIdx.x;
            const int d = i < n ? rs_digit(x[i], dx, Ng, shift) : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            if (d >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[d], __popc(peers));
        }
        __syncthreads();
        if (threadIdx.x < 256) H[(long long)threadIdx.x * ntiles + tile] = hist[threadIdx.x];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(RS_T) rsort_scatter_k(const double* __restrict__ x, const double* __restrict__ u,
                                                        long long n, double dx, int Ng, int shift, int ntiles,
                                                        const int32_t* __restrict__ H, double* __restrict__ xo,
                                                        double* __restrict__ uo, const int32_t* __restrict__ oid,
                                                        int32_t* __restrict__ oido, long long oid_base) {
    extern __shared__ int rs_sm[];
    int (*wcnt)[256] = (int (*)[256])rs_sm;                // elements of warp w with digit d in the current round (kept zero between rounds)
    int (*wpre)[256] = (int (*)[256])(rs_sm + 32 * 256);   // output position of the first of them
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 256; i += RS_T) rs_sm[i] = 0;
    __syncthreads();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // thread d < 256 owns digit d: where the tile's run of that digit starts in the output
        int run = threadIdx.x < 256 ? H[(long long)threadIdx.x * ntiles + tile] : 0;
        const long long base = (long long)tile * RS_TILE;
#pragma unroll 1
        for (int j = 0; j < RS_PER; ++j) {
            const long long i = base + (long long)j * RS_T + threadIdx.x;
            double X = 0.;
            int d = -1;
            if (i < n) { X = x[i]; d = rs_digit(X, dx, Ng, shift); }
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int below = __popc(peers & ((1u << lane) - 1u));
            if (d >= 0 && below == 0) wcnt[w][d] = __popc(peers);
            __syncthreads();
            if (threadIdx.x < 256) {
#pragma unroll 8
                for (int ww = 0; ww < 32; ++ww) {
                    const int c = wcnt[ww][threadIdx.x];
                    wcnt[ww][threadIdx.x] = 0;
                    wpre[ww][threadIdx.x] = run;
                    run += c;
                }
            }
            __syncthreads();
            if (d >= 0) {
                const long long pos = (long long)wpre[w][d] + below;
                xo[pos] = X;
                uo[pos] = u[i];
                // original-index payload (oid == nullptr: this pass starts from the identity numbering)
                if (oido) oido[pos] = oid ? oid[i] : (int32_t)(oid_base + i);
            }
            // the next round rewrites wpre only after its own first barrier, which every thread
            // reaches after the reads above
        }
    }
}
template <bool PERM>
__global__ void __launch_bounds__(SORTB_T) dd_sort_scatter_big_k(DDK k, const double* __restrict__ x0,
                                                                 const double* __restrict__ u0,
                                                                 const double* __restrict__ v0,
                                                                 const double* __restrict__ w0, double* __restrict__ xs,
                                                                 double* __restrict__ us, double* __restrict__ vs,
                                                                 double* __restrict__ ws, int32_t* __restrict__ cursor,
                                                                 const int32_t* __restrict__ oid, int32_t* __restrict__ oids) {
    long long end;
    const long long beg = sortb_range(k.N, &end);
    const unsigned lane = threadIdx.x & 31;
    for (long long base = beg; base < end; base += SORTB_T * SORTB_PER) {
        double X[SORTB_PER], U[SORTB_PER], V[SORTB_PER], W[SORTB_PER];
        int key[SORTB_PER], res[SORTB_PER], O[SORTB_PER];
        unsigned peers[SORTB_PER];
#pragma unroll
        for (int j = 0; j < SORTB_PER; ++j) {
            const long long i = base + j * SORTB_T + threadIdx.x;
            X[j] = 0.; U[j] = 0.; V[j] = 0.; W[j] = 0.; O[j] = (int)i;
            if (i < end) {          // every payload is loaded up front: a load behind the reservation atomics would serialise
                X[j] = x0[i]; if (!PERM) U[j] = u0[i]; if (oid) O[j] = oid[i];
                if (vs) V[j] = v0[i];
                if (ws) W[j] = w0[i];
            }
        }
        // one reservation per distinct key per warp; the four atomics of a thread are independent
#pragma unroll
        for (int j = 0; j < SORTB_PER; ++j) {
            const long long i = base + j * SORTB_T + threadIdx.x;
            key[j] = i < end ? dd_sort_key(k, X[j], i) : -1;
            peers[j] = __match_any_sync(0xffffffffu, key[j]);
            res[j] = 0;
            if (key[j] >= 0 && (int)lane == __ffs(peers[j]) - 1) res[j] = atomicAdd(&cursor[key[j]], __popc(peers[j]));
        }
#pragma unroll
        for (int j = 0; j < SORTB_PER; ++j) {
            const long long i = base + j * SORTB_T + threadIdx.x;
            const int b = __shfl_sync(0xffffffffu, res[j], __ffs(peers[j]) - 1);
            if (key[j] >= 0) {
                const long long pos = (long long)b + __popc(peers[j] & ((1u << lane) - 1u));
                xs[pos] = X[j];
                if (PERM) ((int32_t*)us)[pos] = (int32_t)i; else us[pos] = U[j];
                if (vs) vs[pos] = V[j];
                if (ws) ws[pos] = W[j];
                if (oids) oids[pos] = O[j];
            }
        }
    }
}


// ---- re-injection / bookkeeping of a cell-sorted store that keeps the reference's particle
// numbering: `orig` is the original index of the particle in each slot (the sort's payload); the
// passive velocities v,w are never streamed, so they stay in ORIGINAL order (PIC_L_DD.py:482-483
// copies them unchanged; only re-injection and the thermostat write them) ----
__global__ void dd_apply_draws2_k(const int32_t* __restrict__ slot, const int32_t* __restrict__ orig,
                                  const double* __restrict__ xd, const double* __restrict__ ud,
                                  const double* __restrict__ vd, const double* __restrict__ wd, long long n,
                                  double* __restrict__ x0, double* __restrict__ u0, double* __restrict__ v0,
                                  double* __restrict__ w0, int8_t* __restrict__ active, double* __restrict__ corr) {
    double c1 = 0.0, c2 = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int s = slot[t], o = orig ? orig[t] : s;
        if (xd) x0[s] = xd[t];
        const double un = ud[t];
        if (corr) { const double uo = u0[s]; c1 += un - uo; c2 += un * un - uo * uo; }
        u0[s] = un;
        if (v0) v0[o] = vd[t];
        if (w0) w0[o] = wd[t];
        if (active) active[s] = 1;
    }
    if (corr) {          // corr[0] += sum(u_new - u_old), corr[1] += sum(u_new^2 - u_old^2): turns the moments of the
                         // re-injected state (first Picard iteration) back into those of the state before re-injection
        c1 = warp_sum(c1); c2 = warp_sum(c2);
        if ((threadIdx.x & 31) == 0 && (c1 != 0.0 || c2 != 0.0)) { atomicAdd(corr, c1); atomicAdd(corr + 1, c2); }
    }
}
__global__ void gather_i32_k(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int32_t* __restrict__ dst, long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) dst[t] = src[idx[t]];
}
// dst[idx[t]] = src[t]  (back to the original order); idx == NULL: plain copy
template <typename T>
__global__ void scatter_k(const T* __restrict__ src, const int32_t* __restrict__ idx, T* __restrict__ dst, long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        dst[idx ? idx[t] : t] = src[t];
}
__global__ void invert_perm_k(const int32_t* __restrict__ perm, int32_t* __restrict__ inv, long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) inv[perm[t]] = (int32_t)t;
}
// out[0] += sum u, out[1] += sum u*u  (np.std(u0) and the kinetic-energy diagnostic from ONE pass)
#define MOM_MAX_CTAS 2048
__device__ double g_mom_part[2 * MOM_MAX_CTAS];
__device__ unsigned g_mom_ticket;
__global__ void moments_k(const double* __restrict__ u, long long N, double* __restrict__ out) {
    __shared__ double scratch[33];
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double v = ld_stream(u + i);
        s1 += v; s2 += v * v;
    }
    s1 = block_reduce<0>(s1, scratch);
    s2 = block_reduce<0>(s2, scratch);
    // per-CTA partial sums, added by the CTA that finishes last IN CTA ORDER: for a given particle order the two
    // sums do not depend on scheduling (the reproducible build's diagnostics are bit-reproducible too)
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        g_mom_part[2 * blockIdx.x] = s1; g_mom_part[2 * blockIdx.x + 1] = s2;
        __threadfence();
        const unsigned t = atomicAdd(&g_mom_ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) g_mom_ticket = 0;
        __threadfence();
    }
    __syncthreads();
    if (s_last && threadIdx.x < 2) {
        double a = 0.0;
        for (int c = 0; c < (int)gridDim.x; ++c) a += __ldcg(&g_mom_part[2 * c + threadIdx.x]);
        out[threadIdx.x] += a;
    }
}
// slots named in the absorption log -> re-injected with Philox draws (no flag scan)
__global__ void dd_reinject_philox_log_k(DDK k, const int* __restrict__ buf, double* __restrict__ x0,
                                         double* __restrict__ u0, double* __restrict__ v0, double* __restrict__ w0,
                                         int8_t* __restrict__ active, double s0, double s1, uint64_t seed, uint64_t step,
                                         long long goff, const int32_t* __restrict__ oid) {
    const int n = buf[0];
    if (n > k.dead_cap) return;          // overflow: the flag scan (dd_reinject_philox_k) takes over
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        dd_reinject_one(k, (long long)buf[4 + 4 * t], x0, u0, v0, w0, active, s0, s1, seed, step, goff, oid);
}
// Start of a sheath timestep in ONE launch (PIC_L_DD.py:429-450 + the per-step clears): the re-injection of
// the slots the previous step absorbed -- device Philox draws for the slots named in the absorption log (flag
// scan if the log overflowed), or host draws applied by block 0, which also accumulates the moments'
// correction -- then Es = E0 and the per-step accumulators / statistics / loop flag cleared, and the header
// of the log the coming step writes (another buffer than the one read here) reset.  The host used to
// issue 4-7 tiny operations here with the GPU idle in between (profiles/r2_step_gaps.txt).
struct DDPro {
    const int* log; int* next_log; int philox;
    double s0, s1; uint64_t seed, step; long long goff;
    const int32_t* slot; const int32_t* dorig; const double* xd; const double* ud; const double* vd; const double* wd;
    long long n_draws; double* corr;
    double* x0; double* u0; double* v0; double* w0; int8_t* active; const int32_t* oid;
    double* Es; const double* E0; double* wall_cum; double* stats; long long nstats; int* ctl;
};
__global__ void __launch_bounds__(256) dd_step_prologue_k(DDK k, const DDPro a) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < k.Ng; i += nth) a.Es[i] = a.E0[i];
    if (blockIdx.x == 0) {
        for (long long i = threadIdx.x; i < a.nstats; i += blockDim.x) a.stats[i] = 0.0;
        if (threadIdx.x < 4) { a.wall_cum[threadIdx.x] = 0.0; if (a.next_log) a.next_log[threadIdx.x] = 0; }
        if (threadIdx.x == 0) *a.ctl = 0;
        if (a.corr && threadIdx.x < 2) a.corr[threadIdx.x] = 0.0;
        if (a.n_draws > 0) {
            __syncthreads();
            double c1 = 0.0, c2 = 0.0;
            for (long long t = threadIdx.x; t < a.n_draws; t += blockDim.x) {
                const int s = a.slot[t], o = a.dorig ? a.dorig[t] : s;
                if (a.xd) a.x0[s] = a.xd[t];
                const double un = a.ud[t];
                if (a.corr) { const double uo = a.u0[s]; c1 += un - uo; c2 += un * un - uo * uo; }
                a.u0[s] = un;
                if (a.v0) a.v0[o] = a.vd[t];
                if (a.w0) a.w0[o] = a.wd[t];
                a.active[s] = 1;
            }
            if (a.corr) {
                c1 = warp_sum(c1); c2 = warp_sum(c2);
                if ((threadIdx.x & 31) == 0 && (c1 != 0.0 || c2 != 0.0)) { atomicAdd(a.corr, c1); atomicAdd(a.corr + 1, c2); }
            }
        }
    }
    if (a.philox) {
        const int n = a.log[0];
        if (n <= k.dead_cap) {
            for (int t = tid; t < n; t += nth)
                dd_reinject_one(k, (long long)a.log[4 + 4 * t], a.x0, a.u0, a.v0, a.w0, a.active, a.s0, a.s1, a.seed, a.step, a.goff, a.oid);
        } else {
            // the log overflowed (a step that absorbed more than dead_cap particles): scan the flags
            for (long long i = tid; i < k.N; i += nth)
                if (a.active[i] != 1) dd_reinject_one(k, i, a.x0, a.u0, a.v0, a.w0, a.active, a.s0, a.s1, a.seed, a.step, a.goff, a.oid);
        }
    }
}
// Thermostat, device mode (PIC_L_DD.py:419-427): every ACTIVE particle redraws u,v,w from the ION
// temperature (as written: sqrt(kBTi/m[i]) for both species) with probability gamma.  Philox keyed
// by (seed, step, global original index), so the outcome does not depend on the sort or the sharding.
__global__ void dd_thermostat_philox_k(DDK k, double* __restrict__ u0, double* __restrict__ v0, double* __restrict__ w0,
                                       const int8_t* __restrict__ active, const int32_t* __restrict__ orig, double gamma,
                                       double s0, double s1, uint64_t seed, uint64_t step, long long goff) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        if (active[i] != 1) continue;
        const long long o = orig ? orig[i] : i;
        const uint64_t gid = (uint64_t)(goff + o);
        uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ 0x20000000u};
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        if (!(u52(c[0], c[1]) < gamma)) continue;
        uint32_t d[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ 0x10000000u};
        philox4x32(d, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double twopi = 6.283185307179586;
        const double r1 = sqrt(-2.0 * log(u52(c[2], c[3]))), t1 = twopi * u52(d[0], d[1]);
        const double r2 = sqrt(-2.0 * log(u52(d[2], d[3])));
        uint32_t g[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ 0x08000000u};
        philox4x32(g, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double t2 = twopi * u52(g[0], g[1]);
        const double sg = (i >= k.n_split) ? s1 : s0;
        u0[i] = sg * r1 * cos(t1);
        if (v0) v0[o] = sg * r1 * sin(t1);
        if (w0) w0[o] = sg * r2 * cos(t2);
    }
}

}  // namespace pic

using namespace pic;

// a zeroed slice counter for one launch on stream st (see g_sched_pool)
static int next_sched_slot(cudaStream_t st, int** out) {
    static int* base = nullptr;
    static unsigned next = 0;
    if (!base) PIC_CHECK_CUDA(cudaGetSymbolAddress((void**)&base, g_sched_pool));
    int* p = base + (next++ % PIC_SCHED_SLOTS);
    PIC_CHECK_CUDA(cudaMemsetAsync(p, 0, sizeof(int), st));
    *out = p;
    return PIC_OK;
}

// counting sort driver shared by pic_dev_dd_sort_by_cell / pic_dev_sort_perm_by_cell
template <bool PERM>
static int sort_by_cell_impl(const DDK& k, const double* x0, const double* u0, const double* v0, const double* w0,
                             double* x0s, double* u0s, double* v0s, double* w0s, int32_t* counts, cudaStream_t st,
                             const int32_t* oid = nullptr, int32_t* oids = nullptr) {
    const int nk = 2 * k.Ng;
    const size_t smem = (size_t)nk * sizeof(int);
    const size_t smem_sc = 2 * smem;
    const bool big = smem_sc > (size_t)max_optin_smem() - 1024 || (k.flags & 32);   // bit5: force (experiments)
    // big: counts needs nk + 2 + ceil(nk/1024) + 2 entries (block sums of the three-pass scan behind the keys)
    const int nblk = (nk + 1023) / 1024;
    PIC_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)(nk + 2 + (big ? nblk + 2 : 0)) * sizeof(int32_t), st));
    if (k.N == 0) return PIC_OK;
    if (big) {
        PIC_REQUIRE(nblk <= 1024 * 1024, "sort_by_cell: grid too large");
        int32_t* sums = counts + nk + 2;
        dd_sort_hist_big_k<<<grid_for(k.N, SORTB_T * SORTB_PER, 4), SORTB_T, 0, st>>>(k, x0, counts);
        PIC_CHECK_LAUNCH();
        scan_local_k<<<nblk, 1024, 0, st>>>(counts, nk, sums);
        PIC_CHECK_LAUNCH();
        dd_sort_scan_k<<<1, 1024, 0, st>>>(sums, nblk);
        PIC_CHECK_LAUNCH();
        scan_add_k<<<nblk, 1024, 0, st>>>(counts, nk, sums);
        PIC_CHECK_LAUNCH();
        dd_sort_scatter_big_k<PERM><<<grid_for(k.N, SORTB_T * SORTB_PER, 4), SORTB_T, 0, st>>>(k, x0, u0, v0, w0, x0s, u0s, v0s, w0s, counts, oid, oids);
        PIC_CHECK_LAUNCH();
        return PIC_OK;
    }
    PIC_CHECK_CUDA(cudaFuncSetAttribute(dd_sort_hist_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PIC_CHECK_CUDA(cudaFuncSetAttribute(dd_sort_scatter_k<PERM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sc));
    dd_sort_hist_k<<<grid_for(k.N, 1024, 2), 1024, smem, st>>>(k, x0, counts);
    PIC_CHECK_LAUNCH();
    dd_sort_scan_k<<<1, 1024, 0, st>>>(counts, nk);
    PIC_CHECK_LAUNCH();
    dd_sort_scatter_k<PERM><<<grid_for((k.N + SORT_PER - 1) / SORT_PER, SORT_T, 2), SORT_T, smem_sc, st>>>(
        k, x0, u0, v0, w0, x0s, u0s, v0s, w0s, counts, oid, oids);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

template <bool FIRST, bool TILE, bool AGG>
static int launch_iter(const DDK& k, const double* x0, const double* u0, const double* x1i, double* x1, double* u1,
                       int8_t* active, const double* Es, double* acc, int* range_err, cudaStream_t st) {
    // a short tail is cheaper with the grids left in L2 than with 3*Ng doubles staged per CTA
    if (TILE && (k.N < 16 * (long long)k.Ng || k.fix))      // reproducible build: no shared-memory fp64 atomics
        return launch_iter<FIRST, false, AGG>(k, x0, u0, x1i, x1, u1, active, Es, acc, range_err, st);
    size_t smem = TILE ? (size_t)3 * k.Ng * sizeof(double) : 0;
    auto kern = dd_picard_iter_k<FIRST, TILE, AGG>;
    int per_sm = 8;
    if (TILE) {
        PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        PIC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        per_sm = occ > 0 ? occ : 1;
    }
    kern<<<grid_for(k.N, 256, per_sm), 256, smem, st>>>(k, x0, u0, x1i, x1, u1, active, Es, acc, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

template <bool FIRST>
static int launch_tail(const DDK& k, long long done, const double* x0, const double* u0, const double* x1i, double* x1,
                       double* u1, int8_t* active, const double* Es, double* acc, int* range_err, cudaStream_t st) {
    if (done >= k.N) return PIC_OK;
    DDK t = k;
    t.N = k.N - done;
    t.n_split = k.n_split - done < 0 ? 0 : (k.n_split - done > t.N ? t.N : k.n_split - done);
    t.slot0 = k.slot0 + done;
    return launch_iter<FIRST, true, true>(t, x0 + done, u0 + done, x1i + done, x1 + done, u1 ? u1 + done : nullptr,
                                          active + done, Es, acc, range_err, st);
}

// u1 = u0 + dt*(q/m)*E(xs) for every particle alive at the entry of the LAST Picard iteration,
// 0 for the ones absorbed before it -- the velocity half of the commit when that iteration ran
// without streaming u1 (pic_dev_dd_picard_iter2 with u1 == NULL).
__global__ void __launch_bounds__(256) dd_commit_u_k(DDK k, const double* __restrict__ x0, const double* __restrict__ u0,
                                                     const double* __restrict__ x1_prev, const double* __restrict__ x1_last,
                                                     const int8_t* __restrict__ active, const double* __restrict__ Es,
                                                     double* __restrict__ u1, int first, int tile,
                                                     double* __restrict__ j1_acc, int* __restrict__ range_err) {
    extern __shared__ double sFt[];
    const int Ng = k.Ng;
    const double* sF = Es;                       // large grids: gather from global memory / L2
    if (tile) {
        for (int i = threadIdx.x; i < Ng; i += blockDim.x) sFt[i] = Es[i];
        __syncthreads();
        sF = sFt;
    }
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k.N; i += (long long)gridDim.x * blockDim.x) {
        const double X0 = ld_stream(x0 + i), U0 = ld_stream(u0 + i);
        double xs = X0;
        if (!first) {
            // absorbed before the last iteration <=> flagged and the last iteration wrote the
            // reference's zero instead of a position (PIC_L_DD.py:459-462)
            if (active[i] != 1 && x1_last[i] == 0.0) { st_stream(u1 + i, 0.0); continue; }
            xs = (X0 + ld_stream(x1_prev + i)) * 0.5;
        }
        Cell c = cell_dd_fast(xs, k.dx, k.idx);
        if (c.iL < 0 || c.iL > Ng - 2) { ++bad; c.iL = clampi(c.iL, 0, Ng - 2); }
        const double Ei = c.wL * sF[c.iL] + c.wR * sF[c.iL + 1];
        const bool sp = i >= k.n_split;
        const double U1 = U0 + (sp ? k.c1[1] : k.c1[0]) * Ei;
        st_stream(u1 + i, U1);
        // j1 of a light last iteration: the survivors' current at n+1 (PIC_L_DD.py:513)
        if (j1_acc && active[i] == 1) {
            Cell b = cell_dd(x1_last[i], k.dx);
            if (b.iL < 0 || b.iL > Ng - 2) { ++bad; b.iL = clampi(b.iL, 0, Ng - 2); }
            const double qf = (sp ? k.q[1] : k.q[0]) * U1 * k.p2c;
            acc_add(k, j1_acc - Ng, Ng + b.iL, qf * b.wL * k.idx); acc_add(k, j1_acc - Ng, Ng + b.iL + 1, qf * b.wR * k.idx);
        }
    }
    if (bad && range_err) atomicAdd(range_err, bad);
}

// j1 part of the field phase alone (PIC_L_DD.py:55-66 applied to j1, :551): wall terms from the
// cumulative counts, edge fold, mean; the j1 accumulator is zeroed.  One CTA.
__global__ void __launch_bounds__(1024) dd_j1_finish_k(DDK k, double* __restrict__ acc, const double* __restrict__ wall_cum,
                                                       double* __restrict__ j1o, double* __restrict__ stats) {
    __shared__ double scratch[33];
    const int Ng = k.Ng;
    const double wallL = wall_cum[0] * (k.dx * k.q[0] * k.p2c / k.dt) + wall_cum[1] * (k.dx * k.q[1] * k.p2c / k.dt);
    const double wallR = wall_cum[2] * (-k.dx * k.q[0] * k.p2c / k.dt) + wall_cum[3] * (-k.dx * k.q[1] * k.p2c / k.dt);
    double* j1 = acc + Ng;
    if (k.fix) {
        for (int i = threadIdx.x; i < Ng; i += blockDim.x) j1[i] += fix_take(k, Ng + i);
        __syncthreads();
    }
    double s1 = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        double b = j1[i];
        if (i == 0) b = (b + wallL) + j1[1];
        if (i == Ng - 1) b = (b + wallR) + j1[Ng - 2];
        s1 += b;
        j1o[i] = b;
    }
    s1 = block_reduce<0>(s1, scratch);
    __syncthreads();
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) j1[i] = 0.0;
    if (threadIdx.x == 0) stats[1] = s1 / (double)Ng;
}

extern "C" {

int pic_dev_dd_picard_iter2(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_in,
                            double* x1_out, double* u1, int8_t* active, const double* Es, double* acc, int first,
                            int* range_err, void* stream) {
    return pic_dev_dd_picard_iter3(p, x0, u0, x1_in, x1_out, u1, active, Es, acc, first, range_err, nullptr, stream);
}

int pic_dev_dd_picard_iter3(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_in,
                            double* x1_out, double* u1, int8_t* active, const double* Es, double* acc, int first,
                            int* range_err, const int32_t* done, void* stream) {
    return pic_dev_dd_picard_iter4(p, x0, u0, x1_in, x1_out, u1, active, Es, acc, first, range_err, done, nullptr, 0, nullptr,
                                   0, stream);
}

int pic_dev_dd_picard_iter4(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_in,
                            double* x1_out, double* u1, int8_t* active, const double* Es, double* acc, int first,
                            int* range_err, const int32_t* done, int32_t* dead_buf, int32_t dead_cap,
                            const int32_t* orig, int32_t iteration, void* stream) {
    return pic_dev_dd_picard_iter5(p, x0, u0, x1_in, x1_out, u1, active, Es, acc, first, range_err, done, dead_buf, dead_cap, orig,
                                   iteration, nullptr, stream);
}

int pic_dev_dd_picard_iter5(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_in,
                            double* x1_out, double* u1, int8_t* active, const double* Es, double* acc, int first,
                            int* range_err, const int32_t* done, int32_t* dead_buf, int32_t dead_cap,
                            const int32_t* orig, int32_t iteration, double* moments, void* stream) {
    PIC_REQUIRE(p && x0 && u0 && x1_in && x1_out && active && Es && acc, "dd_picard_iter: null pointer");
    PIC_REQUIRE(!moments || first, "dd_picard_iter: the velocity moments are accumulated by the FIRST iteration only");
    PIC_REQUIRE(p->N >= 0 && p->Ng >= 3 && p->dx > 0 && p->dt > 0, "dd_picard_iter: bad parameters");
    PIC_REQUIRE(!dead_buf || (dead_cap > 0 && ((uintptr_t)dead_buf & 15) == 0), "dd_picard_iter: absorption log without storage / unaligned");
    if (p->N == 0) return PIC_OK;
    DDK k = make_ddk(p);
    k.dead_buf = dead_buf; k.oid = orig; k.dead_cap = dead_cap; k.iter = iteration; k.mom = moments;
    PIC_REQUIRE(!(p->flags & 128) || !(p->flags & (1 | 2 | 4 | 8)),
                "dd_picard_iter: the reproducible build (flags bit7) exists for the default window kernel only");
    ddk_bind_fix(k, acc, range_err);
    k.done = done;
    cudaStream_t st = (cudaStream_t)stream;
    const double* x1i = x1_in;
    double* x1 = x1_out;
    bool tile = !(p->flags & 2) && (size_t)3 * k.Ng * sizeof(double) <= (size_t)max_optin_smem() - 1024;
    bool agg = !(p->flags & 1);
    size_t smem5 = ((size_t)3 * k.Ng + (size_t)2 * V5_W * V5_T) * sizeof(double);
    const size_t smem6 = ((size_t)((k.Ng + 15) & ~15) + (size_t)2 * V6_W * V6_T + (size_t)(V6_T / 32) * V6_NST * 192 +
                          (size_t)(V6_T / 32) * V6_NST) * sizeof(double);
    const bool aligned16 = (((uintptr_t)x0 | (uintptr_t)u0 | (uintptr_t)x1i | (uintptr_t)x1 | (uintptr_t)u1) & 15) == 0;
    // large-grid build of the same kernel: per-warp field windows instead of the whole-grid tile, so
    // the shared-memory footprint does not depend on Ng (flags bit4 forces it, for tests)
    const size_t smem6b = ((size_t)(V6_T / 32) * V6_EW + (size_t)2 * V6_W * V6_T + (size_t)(V6_T / 32) * V6_NST * 192 +
                           (size_t)(V6_T / 32) * V6_NST) * sizeof(double);
    const bool big = ((p->flags & 16) || smem6 > (size_t)max_optin_smem() - 512) && k.Ng >= V6_EW;
    if (big && aligned16 && !(p->flags & (1 | 4 | 8))) {
        const long long nchunks = k.N / V6_CHUNK;
        if (nchunks > 0) {
            // the wide-window build (15-node deposit windows) holds more cells per window, so the windows are
            // flushed and re-centred less often (env PIC_V6_NARROW=1 keeps the 7-node build, for A/B runs)
            static const bool narrow_b = [] { const char* e = getenv("PIC_V6_NARROW"); return e && e[0] == '1'; }();
            const bool wide = !narrow_b;
            const size_t smem6bw = ((size_t)(V6_T / 32) * V6_EW + (size_t)2 * V6_W_WIDE * V6_T + (size_t)(V6_T / 32) * V6_NST_WIDE * 192 +
                                    (size_t)(V6_T / 32) * V6_NST_WIDE) * sizeof(double);
            auto kern = wide ? (first ? (u1 ? dd_picard_iter_v6_k<true, true, true, true> : dd_picard_iter_v6_k<true, false, true, true>)
                                      : (u1 ? dd_picard_iter_v6_k<false, true, true, true> : dd_picard_iter_v6_k<false, false, true, true>))
                             : (first ? (u1 ? dd_picard_iter_v6_k<true, true, true> : dd_picard_iter_v6_k<true, false, true>)
                                      : (u1 ? dd_picard_iter_v6_k<false, true, true> : dd_picard_iter_v6_k<false, false, true>));
            const size_t smemb = wide ? smem6bw : smem6b;
            PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemb));
            long long cap = device_sm_count();
            int grid = (int)(nchunks < cap ? nchunks : cap);
            // rows (of 64 particles) per deposit/field window: about three cells' worth of particles with 7-node
            // windows, eight with 15-node ones (the rest of the window is for the drift between sorts)
            const double ppc = (double)k.N / 2.0 / (double)k.Ng;
            int fr = 16;
            while (fr > 1 && 64.0 * fr > (wide ? 8.0 : 3.0) * ppc) fr >>= 1;
            PIC_REQUIRE(nchunks < (1 << 28), "dd_picard_iter: shard too large");
            int* sched = nullptr;
            int rcs = next_sched_slot(st, &sched);
            if (rcs) return rcs;
            kern<<<grid, V6_T, smemb, st>>>(k, (int)((unsigned)nchunks | ((unsigned)(fr - 1) << 28)), x0, u0, x1i, x1, u1, active, Es, acc, range_err, sched);
            PIC_CHECK_LAUNCH();
        }
        const long long done = nchunks * V6_CHUNK;
        if (done >= k.N || nchunks > 0) return PIC_OK;      // the kernel finishes the ragged tail itself
        DDK t = k;
        t.N = k.N - done;
        t.n_split = k.n_split - done < 0 ? 0 : (k.n_split - done > t.N ? t.N : k.n_split - done);
        t.slot0 = k.slot0 + done;
        // tail: grid-stride kernel with the grids left in global memory
        return first ? launch_iter<true, false, true>(t, x0 + done, u0 + done, x1i + done, x1 + done, u1 ? u1 + done : nullptr,
                                                      active + done, Es, acc, range_err, st)
                     : launch_iter<false, false, true>(t, x0 + done, u0 + done, x1i + done, x1 + done, u1 ? u1 + done : nullptr,
                                                       active + done, Es, acc, range_err, st);
    }
    if (!(p->flags & (1 | 2 | 4 | 8)) && aligned16 && smem6 <= (size_t)max_optin_smem() - 512) {
        // default: TMA-staged private-window kernel, one persistent CTA per SM; the wide-window build when the
        // field tile leaves room for it (env PIC_V6_NARROW=1 keeps the 7-node build, for A/B runs)
        const long long nchunks = k.N / V6_CHUNK;
        if (nchunks > 0) {
            const size_t smem6w = ((size_t)((k.Ng + 15) & ~15) + (size_t)2 * V6_W_WIDE * V6_T + (size_t)(V6_T / 32) * V6_NST_WIDE * 192 +
                                   (size_t)(V6_T / 32) * V6_NST_WIDE) * sizeof(double);
            static const bool narrow = [] { const char* e = getenv("PIC_V6_NARROW"); return e && e[0] == '1'; }();
            const bool wide = !narrow && smem6w <= (size_t)max_optin_smem() - 512;
            auto kern = wide ? (first ? (u1 ? dd_picard_iter_v6_k<true, true, false, true> : dd_picard_iter_v6_k<true, false, false, true>)
                                      : (u1 ? dd_picard_iter_v6_k<false, true, false, true> : dd_picard_iter_v6_k<false, false, false, true>))
                             : (first ? (u1 ? dd_picard_iter_v6_k<true, true> : dd_picard_iter_v6_k<true, false>)
                                      : (u1 ? dd_picard_iter_v6_k<false, true> : dd_picard_iter_v6_k<false, false>));
            const size_t smem = wide ? smem6w : smem6;
            PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            long long cap = device_sm_count();
            int grid = (int)(nchunks < cap ? nchunks : cap);
            int* sched = nullptr;
            int rcs = next_sched_slot(st, &sched);
            if (rcs) return rcs;
            kern<<<grid, V6_T, smem, st>>>(k, (int)nchunks, x0, u0, x1i, x1, u1, active, Es, acc, range_err, sched);
            PIC_CHECK_LAUNCH();
            return PIC_OK;                                   // the kernel finishes the ragged tail itself
        }
        const long long done = nchunks * V6_CHUNK;
        return first ? launch_tail<true>(k, done, x0, u0, x1i, x1, u1, active, Es, acc, range_err, st)
                     : launch_tail<false>(k, done, x0, u0, x1i, x1, u1, active, Es, acc, range_err, st);
    }
    if (!(p->flags & (1 | 2 | 4)) && smem5 <= (size_t)max_optin_smem() - 512) {
        // register-prefetch build of the window kernel (kept for comparison)
        const long long nchunks = k.N / V5_CHUNK;
        if (nchunks > 0) {
            auto kern = first ? dd_picard_iter_v5_k<true> : dd_picard_iter_v5_k<false>;
            PIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));
            long long cap = device_sm_count();
            int grid = (int)(nchunks < cap ? nchunks : cap);
            kern<<<grid, V5_T, smem5, st>>>(k, nchunks, x0, u0, x1i, x1, u1, active, Es, acc, range_err);
            PIC_CHECK_LAUNCH();
        }
        const long long done = nchunks * V5_CHUNK;
        return first ? launch_tail<true>(k, done, x0, u0, x1i, x1, u1, active, Es, acc, range_err, st)
                     : launch_tail<false>(k, done, x0, u0, x1i, x1, u1, active, Es, acc, range_err, st);
    }
#define PIC_DD_DISPATCH(F, T, A) return launch_iter<F, T, A>(k, x0, u0, x1i, x1, u1, active, Es, acc, range_err, st)
    if (first) {
        if (tile) { if (agg) PIC_DD_DISPATCH(true, true, true); else PIC_DD_DISPATCH(true, true, false); }
        else { if (agg) PIC_DD_DISPATCH(true, false, true); else PIC_DD_DISPATCH(true, false, false); }
    } else {
        if (tile) { if (agg) PIC_DD_DISPATCH(false, true, true); else PIC_DD_DISPATCH(false, true, false); }
        else { if (agg) PIC_DD_DISPATCH(false, false, true); else PIC_DD_DISPATCH(false, false, false); }
    }
#undef PIC_DD_DISPATCH
}

int pic_dev_dd_picard_iter(const pic_dd_params* p, const double* x0, const double* u0, double* x1, double* u1,
                           int8_t* active, const double* Es, double* acc, int first, int* range_err, void* stream) {
    PIC_REQUIRE(u1, "dd_picard_iter: null pointer");
    return pic_dev_dd_picard_iter2(p, x0, u0, x1, x1, u1, active, Es, acc, first, range_err, stream);
}

int pic_dev_dd_commit_u(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_prev,
                        const double* x1_last, const int8_t* active, const double* Es, double* u1, int first,
                        int* range_err, void* stream) {
    return pic_dev_dd_commit_u2(p, x0, u0, x1_prev, x1_last, active, Es, u1, first, nullptr, range_err, stream);
}

int pic_dev_dd_j1_finish(const pic_dd_params* p, double* acc, const double* wall_cum, double* j1, double* stats,
                         void* stream) {
    PIC_REQUIRE(p && acc && wall_cum && j1 && stats, "dd_j1_finish: null pointer");
    DDK k = make_ddk(p);
    ddk_bind_fix(k, acc, nullptr);
    dd_j1_finish_k<<<1, 1024, 0, (cudaStream_t)stream>>>(k, acc, wall_cum, j1, stats);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_dd_commit_u2(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_prev,
                         const double* x1_last, const int8_t* active, const double* Es, double* u1, int first,
                         double* acc, int* range_err, void* stream) {
    PIC_REQUIRE(p && x0 && u0 && x1_prev && x1_last && active && Es && u1, "dd_commit_u: null pointer");
    if (p->N == 0) return PIC_OK;
    DDK k = make_ddk(p);
    ddk_bind_fix(k, acc, range_err);
    size_t smem = (size_t)k.Ng * sizeof(double);
    const int tile = smem <= (size_t)max_optin_smem() - 1024;
    if (!tile) smem = 0;
    PIC_CHECK_CUDA(cudaFuncSetAttribute(dd_commit_u_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem ? smem : 1)));
    int occ = 0;
    PIC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dd_commit_u_k, 256, smem));
    dd_commit_u_k<<<grid_for(k.N, 256, occ > 0 ? occ : 1), 256, smem, (cudaStream_t)stream>>>(
        k, x0, u0, x1_prev, x1_last, active, Es, u1, first, tile, acc ? acc + k.Ng : nullptr, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_debug_cta_timer(uint64_t* buf) {
    unsigned long long* b = (unsigned long long*)buf;
    PIC_CHECK_CUDA(cudaMemcpyToSymbol(g_cta_timer, &b, sizeof(b)));
    return PIC_OK;
}

int pic_dev_selftest_div(double b, uint64_t n, uint64_t seed, uint64_t* mismatches_dev, void* stream) {
    PIC_REQUIRE(b > 0 && mismatches_dev, "selftest_div: bad argument");
    selftest_div_k<<<device_sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(b, n, seed,
                                                                           (unsigned long long*)mismatches_dev);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_dd_field_update(const pic_dd_params* p, double* acc, double* wall_cum, const double* E0, double* Es,
                            double* E1, double* j1, double* stats, void* stream) {
    return pic_dev_dd_field_update2(p, acc, wall_cum, E0, Es, E1, j1, stats, nullptr, nullptr, nullptr, 0.0, 0, stream);
}

int pic_dev_dd_field_update2(const pic_dd_params* p, double* acc, double* wall_cum, const double* E0, double* Es,
                             double* E1, double* j1, double* stats, double* Es_prev, double* rhist, int32_t* ctl,
                             double tol, int maxiter, void* stream) {
    PIC_REQUIRE(p && acc && wall_cum && E0 && Es && E1 && j1 && stats, "dd_field_update: null pointer");
    PIC_REQUIRE(!(ctl || rhist) || maxiter >= 1, "dd_field_update: maxiter must be >= 1 with ctl / rhist");
    DDK k = make_ddk(p);
    ddk_bind_fix(k, acc, nullptr);
    if (k.Ng > 32768) {
        // stats[4..7] is the reduction scratch of the cooperative kernel (the caller provides 8 doubles)
        double* red = stats + 4;
        int grid = (k.Ng + 1023) / 1024;
        if (grid > device_sm_count()) grid = device_sm_count();
        void* args[] = {&k, &acc, &wall_cum, (void*)&E0, &Es, &E1, &j1, &stats, &red, &Es_prev, &rhist, &ctl, &tol, &maxiter};
        PIC_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)dd_field_update_big_k, dim3(grid), dim3(1024), args, 0,
                                                   (cudaStream_t)stream));
        return PIC_OK;
    }
    dd_field_update_k<<<1, 1024, 0, (cudaStream_t)stream>>>(k, acc, wall_cum, E0, Es, E1, j1, stats, Es_prev, rhist, ctl, tol,
                                                            maxiter);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

// ---- peer-memory reduction (see dd_field_update_p2p_k) ----
int pic_p2p_alloc(int64_t nacc, int world, void** dev_ptr, void* handle64) {
    PIC_REQUIRE(nacc > 0 && world >= 1 && world <= 64 && dev_ptr && handle64, "p2p_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    PIC_REQUIRE(nacc % 2 == 0, "p2p_alloc: nacc must be even (128-bit copies)");
    // acc[nacc] | inbox[2][world][nacc] | ready[2][world] | nreal, cnt  (see dd_field_update_p2p_k)
    const size_t bytes = (size_t)nacc * (1 + 2 * (size_t)world) * sizeof(double) + ((size_t)2 * world + 2) * sizeof(unsigned);
    void* p = nullptr;
    PIC_CHECK_CUDA(cudaMalloc(&p, bytes));
    PIC_CHECK_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    PIC_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, sizeof(h));
    PIC_CHECK_CUDA(cudaDeviceSynchronize());
    *dev_ptr = p;
    return PIC_OK;
}

int pic_p2p_set_timeout(int log2_spins) {
    PIC_REQUIRE(log2_spins >= 10 && log2_spins <= 34, "p2p_set_timeout: log2_spins must be in [10, 34]");
    PIC_CHECK_CUDA(cudaMemcpyToSymbol(g_p2p_spin_log2, &log2_spins, sizeof(int)));
    return PIC_OK;
}

int pic_p2p_open(const void* handle64, void** peer_ptr) {
    PIC_REQUIRE(handle64 && peer_ptr, "p2p_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    PIC_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr = p;
    return PIC_OK;
}

int pic_p2p_close(void* peer_ptr) {
    if (peer_ptr) PIC_CHECK_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return PIC_OK;
}

int pic_p2p_free(void* dev_ptr) {
    if (dev_ptr) PIC_CHECK_CUDA(cudaFree(dev_ptr));
    return PIC_OK;
}

int pic_dev_p2p_reduce(const double* const* peers_dev, int rank, int world, uint32_t seq, int64_t nacc, double* sum,
                       int* err, void* stream) {
    PIC_REQUIRE(peers_dev && sum && err && world >= 1 && world <= 64 && rank >= 0 && rank < world && nacc > 0,
                "p2p_reduce: bad argument");
    (void)seq;                   // the kernels count the reductions themselves
    PIC_REQUIRE(nacc % 2 == 0, "p2p_reduce: nacc must be even");
    P2P P{(double* const*)peers_dev, rank, world, (int)nacc, err};
    p2p_reduce_k<<<P2P_NB, 1024, 0, (cudaStream_t)stream>>>(P, sum);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_dd_field_update_p2p(const pic_dd_params* p, const double* const* peers_dev, int rank, int world, uint32_t seq,
                                double* acc_sum, double* wall_cum, const double* E0, double* Es, double* E1, double* j1,
                                double* stats, double* Es_prev, double* rhist, int32_t* ctl, double tol, int maxiter,
                                int* err, void* stream) {
    PIC_REQUIRE(p && peers_dev && acc_sum && wall_cum && E0 && Es && E1 && j1 && stats && err, "dd_field_update_p2p: null pointer");
    PIC_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "dd_field_update_p2p: bad rank / world");
    PIC_REQUIRE(!(ctl || rhist) || maxiter >= 1, "dd_field_update_p2p: maxiter must be >= 1 with ctl / rhist");
    PIC_REQUIRE(p->Ng <= 32768 && !(p->flags & 128), "dd_field_update_p2p: one-CTA field phase of the default build only");
    DDK k = make_ddk(p);
    (void)seq;                   // the kernels count the reductions themselves
    P2P P{(double* const*)peers_dev, rank, world, 2 * k.Ng + 4, err};
    dd_field_update_p2p_k<<<P2P_NB, 1024, 0, (cudaStream_t)stream>>>(k, P, acc_sum, wall_cum, E0, Es, E1, j1, stats, Es_prev, rhist,
                                                                ctl, tol, maxiter);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_dd_interpolate(const double* F, const double* x, double* out, int64_t N, int Ng, double dx,
                           int* range_err, void* stream) {
    PIC_REQUIRE(F && x && out && N >= 0 && Ng >= 2, "dd_interpolate: bad argument");
    if (N == 0) return PIC_OK;
    dd_interpolate_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(F, x, out, N, Ng, dx, range_err);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

// `out` doubles as the accumulator: it must provide Ng+2 doubles of scratch BEFORE the
// fold, so the host wrappers pass a scratch buffer through pic_dev_dd_weight_raw.
int pic_dev_dd_weight(const double* x, const double* q, const double* v, const double* active, double* out,
                      int64_t N, int Ng, double dx, double dt, double p2c, int* range_err, void* stream) {
    PIC_REQUIRE(x && q && active && out && N >= 0 && Ng >= 3, "dd_weight: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    double* acc = nullptr;
    PIC_CHECK_CUDA(cudaMallocAsync((void**)&acc, (size_t)(Ng + 2) * sizeof(double), st));
    PIC_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)(Ng + 2) * sizeof(double), st));
    if (N > 0) {
        size_t smem = (size_t)(Ng + 2) * sizeof(double);
        if (v) {
            PIC_CHECK_CUDA(cudaFuncSetAttribute(dd_weight_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dd_weight_k<true><<<grid_for(N, 256, 4), 256, smem, st>>>(x, q, v, active, acc, N, Ng, dx, dt, p2c, range_err);
        } else {
            PIC_CHECK_CUDA(cudaFuncSetAttribute(dd_weight_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dd_weight_k<false><<<grid_for(N, 256, 4), 256, smem, st>>>(x, q, v, active, acc, N, Ng, dx, dt, p2c, range_err);
        }
        PIC_CHECK_LAUNCH();
    }
    dd_weight_fold_k<<<1, 1024, 0, st>>>(acc, out, Ng, v != nullptr);
    PIC_CHECK_LAUNCH();
    PIC_CHECK_CUDA(cudaFreeAsync(acc, st));
    return PIC_OK;
}

int pic_dev_dd_apply_draws(const int32_t* idx, const double* xd, const double* ud, const double* vd,
                           const double* wd, int64_t n, double* x0, double* u0, double* v0, double* w0,
                           int8_t* active, void* stream) {
    PIC_REQUIRE(n >= 0, "dd_apply_draws: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(idx && xd && ud && x0 && u0 && active, "dd_apply_draws: null pointer");
    PIC_REQUIRE((!v0 || vd) && (!w0 || wd), "dd_apply_draws: v/w draws missing");
    dd_apply_draws_k<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(idx, xd, ud, vd, wd, n, x0, u0, v0, w0, active);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_dd_reinject_philox(const pic_dd_params* p, double* x0, double* u0, double* v0, double* w0,
                               int8_t* active, const double sigma[2], uint64_t seed, uint64_t step,
                               int64_t global_offset, void* stream) {
    return pic_dev_dd_reinject_philox2(p, x0, u0, v0, w0, active, nullptr, sigma, seed, step, global_offset, nullptr, 0, stream);
}

int pic_dev_dd_reinject_philox2(const pic_dd_params* p, double* x0, double* u0, double* v0, double* w0,
                                int8_t* active, const int32_t* orig, const double sigma[2], uint64_t seed,
                                uint64_t step, int64_t global_offset, const int32_t* dead_count, int32_t dead_cap,
                                void* stream) {
    PIC_REQUIRE(p && x0 && u0 && active && sigma, "dd_reinject_philox: null pointer");
    if (p->N == 0) return PIC_OK;
    DDK k = make_ddk(p);
    k.dead_cap = dead_cap;
    dd_reinject_philox_k<<<grid_for((k.N + 15) / 16, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        k, x0, u0, v0, w0, active, sigma[0], sigma[1], seed, step, global_offset, orig, dead_count);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_sum_sq(const double* u, int64_t N, double scale, double* out1, void* stream) {
    PIC_REQUIRE(u && out1 && N >= 0, "sum_sq: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    PIC_CHECK_CUDA(cudaMemsetAsync(out1, 0, sizeof(double), st));
    if (N == 0) return PIC_OK;
    sum_sq_k<<<grid_for(N, 256, 8), 256, 0, st>>>(u, N, scale, out1);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

static int sort_by_cell_stable_impl(const pic_dd_params* p, double* x0, double* u0, double* xs, double* us, int32_t* o0,
                                    int32_t* o1, int identity, int32_t* scratch, int64_t scratch_entries,
                                    int32_t* result_in_scratch, void* stream) {
    PIC_REQUIRE(p && x0 && u0 && xs && us && scratch && result_in_scratch, "dd_sort_by_cell_stable: null pointer");
    PIC_REQUIRE(p->N >= 0 && p->N < 2147483647LL && p->Ng >= 2 && p->dx > 0, "dd_sort_by_cell_stable: bad parameters");
    PIC_REQUIRE((o0 == nullptr) == (o1 == nullptr), "dd_sort_by_cell_stable: both payload buffers or none");
    cudaStream_t st = (cudaStream_t)stream;
    int bits = 0;
    while ((1 << bits) < p->Ng) ++bits;           // cells 0 .. Ng-1
    const int passes = (bits + 7) / 8;
    *result_in_scratch = passes & 1;
    const long long blk[3] = {0, p->n_split < 0 ? 0 : (p->n_split > p->N ? p->N : p->n_split), p->N};
    for (int sp = 0; sp < 2; ++sp) {
        const long long off = blk[sp], n = blk[sp + 1] - blk[sp];
        if (n <= 0) continue;
        const long long ntl = (n + RS_TILE - 1) / RS_TILE;
        const long long nH = 256 * ntl, nblk = (nH + 1023) / 1024;
        PIC_REQUIRE(nH + nblk + 2 <= scratch_entries && nH < 2147483647LL && nblk <= 1024 * 1024,
                    "dd_sort_by_cell_stable: scratch too small");
        int32_t* H = scratch;
        int32_t* sums = scratch + nH;
        const int ntiles = (int)ntl;
        const int grid = (int)(ntl < (long long)device_sm_count() * 2 ? ntl : (long long)device_sm_count() * 2);
        double *sx = x0 + off, *su = u0 + off, *dx_ = xs + off, *du = us + off;
        int32_t *so = o0 ? o0 + off : nullptr, *dso = o1 ? o1 + off : nullptr;
        const size_t smem = (size_t)2 * 32 * 256 * sizeof(int);
        PIC_CHECK_CUDA(cudaFuncSetAttribute(rsort_scatter_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int ps = 0; ps < passes; ++ps) {
            rsort_hist_k<<<grid, RS_T, 0, st>>>(sx, n, p->dx, p->Ng, 8 * ps, ntiles, H);
            PIC_CHECK_LAUNCH();
            scan_local_k<<<(int)nblk, 1024, 0, st>>>(H, (int)nH, sums);
            PIC_CHECK_LAUNCH();
            dd_sort_scan_k<<<1, 1024, 0, st>>>(sums, (int)nblk);
            PIC_CHECK_LAUNCH();
            scan_add_k<<<(int)nblk, 1024, 0, st>>>(H, (int)nH, sums);
            PIC_CHECK_LAUNCH();
            rsort_scatter_k<<<grid, RS_T, smem, st>>>(sx, su, n, p->dx, p->Ng, 8 * ps, ntiles, H, dx_, du,
                                                      (ps == 0 && identity) ? nullptr : so, dso, off);
            PIC_CHECK_LAUNCH();
            double* t = sx; sx = dx_; dx_ = t;
            t = su; su = du; du = t;
            int32_t* ti = so; so = dso; dso = ti;
        }
    }
    return PIC_OK;
}

int pic_dev_dd_sort_by_cell_stable(const pic_dd_params* p, double* x0, double* u0, double* xs, double* us,
                                   int32_t* scratch, int64_t scratch_entries, int32_t* result_in_scratch,
                                   void* stream) {
    return sort_by_cell_stable_impl(p, x0, u0, xs, us, nullptr, nullptr, 0, scratch, scratch_entries, result_in_scratch, stream);
}

int pic_dev_dd_sort_by_cell_stable2(const pic_dd_params* p, double* x0, double* u0, double* xs, double* us, int32_t* orig,
                                    int32_t* origs, int identity, int32_t* scratch, int64_t scratch_entries,
                                    int32_t* result_in_scratch, void* stream) {
    PIC_REQUIRE(orig && origs, "dd_sort_by_cell_stable2: null payload buffer");
    return sort_by_cell_stable_impl(p, x0, u0, xs, us, orig, origs, identity, scratch, scratch_entries, result_in_scratch, stream);
}

int pic_dev_dd_sort_by_cell(const pic_dd_params* p, const double* x0, const double* u0, const double* v0,
                            const double* w0, double* x0s, double* u0s, double* v0s, double* w0s, int32_t* counts,
                            void* stream) {
    PIC_REQUIRE(p && x0 && u0 && x0s && u0s && counts, "dd_sort_by_cell: null pointer");
    PIC_REQUIRE(p->N < 2147483647LL, "dd_sort_by_cell: shard too large for int32 cursors");
    return sort_by_cell_impl<false>(make_ddk(p), x0, u0, v0, w0, x0s, u0s, v0s, w0s, counts, (cudaStream_t)stream);
}

int pic_dev_sort_perm_by_cell(const pic_dd_params* p, const double* x, double* xs, int32_t* perm, int32_t* counts,
                              void* stream) {
    PIC_REQUIRE(p && x && xs && perm && counts, "sort_perm_by_cell: null pointer");
    PIC_REQUIRE(p->N < 2147483647LL, "sort_perm_by_cell: store too large for int32 indices");
    return sort_by_cell_impl<true>(make_ddk(p), x, nullptr, nullptr, nullptr, xs, (double*)perm, nullptr, nullptr, counts,
                                   (cudaStream_t)stream);
}

int pic_dev_dd_sort_by_cell2(const pic_dd_params* p, const double* x0, const double* u0, const int32_t* orig,
                             double* x0s, double* u0s, int32_t* origs, int32_t* counts, void* stream) {
    PIC_REQUIRE(p && x0 && u0 && x0s && u0s && origs && counts, "dd_sort_by_cell2: null pointer");
    PIC_REQUIRE(p->N < 2147483647LL, "dd_sort_by_cell2: shard too large for int32 cursors");
    return sort_by_cell_impl<false>(make_ddk(p), x0, u0, nullptr, nullptr, x0s, u0s, nullptr, nullptr, counts,
                                    (cudaStream_t)stream, orig, origs);
}

int pic_dev_sort_by_cell_payload(const pic_dd_params* p, const double* x, const double* a, const double* b, const double* c,
                                 double* xs, double* as, double* bs, double* cs, int32_t* perm, int32_t* counts,
                                 void* stream) {
    PIC_REQUIRE(p && x && a && xs && as && counts, "sort_by_cell_payload: null pointer");
    PIC_REQUIRE((!b || bs) && (!c || cs), "sort_by_cell_payload: payload output missing");
    PIC_REQUIRE(p->N < 2147483647LL, "sort_by_cell_payload: store too large for int32 indices");
    return sort_by_cell_impl<false>(make_ddk(p), x, a, b, c, xs, as, b ? bs : nullptr, c ? cs : nullptr, counts,
                                    (cudaStream_t)stream, nullptr, perm);
}

int pic_dev_dd_apply_draws2(const int32_t* slot, const int32_t* orig, const double* xd, const double* ud,
                            const double* vd, const double* wd, int64_t n, double* x0, double* u0, double* v0,
                            double* w0, int8_t* active, void* stream) {
    PIC_REQUIRE(n >= 0, "dd_apply_draws2: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(slot && ud && u0, "dd_apply_draws2: null pointer");
    PIC_REQUIRE((!xd || x0) && (!v0 || vd) && (!w0 || wd), "dd_apply_draws2: draws / targets missing");
    return pic_dev_dd_apply_draws3(slot, orig, xd, ud, vd, wd, n, x0, u0, v0, w0, active, nullptr, stream);
}

int pic_dev_dd_apply_draws3(const int32_t* slot, const int32_t* orig, const double* xd, const double* ud,
                            const double* vd, const double* wd, int64_t n, double* x0, double* u0, double* v0,
                            double* w0, int8_t* active, double* corr, void* stream) {
    PIC_REQUIRE(n >= 0, "dd_apply_draws3: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(slot && ud && u0, "dd_apply_draws3: null pointer");
    PIC_REQUIRE((!xd || x0) && (!v0 || vd) && (!w0 || wd), "dd_apply_draws3: draws / targets missing");
    dd_apply_draws2_k<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(slot, orig, xd, ud, vd, wd, n, x0, u0, v0, w0, active, corr);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_dd_reinject_philox_log(const pic_dd_params* p, const int32_t* dead_buf, int32_t dead_cap, double* x0,
                                   double* u0, double* v0, double* w0, int8_t* active, const int32_t* orig,
                                   const double sigma[2], uint64_t seed, uint64_t step, int64_t global_offset,
                                   void* stream) {
    PIC_REQUIRE(p && dead_buf && dead_cap > 0 && x0 && u0 && active && sigma, "dd_reinject_philox_log: null pointer");
    if (p->N == 0) return PIC_OK;
    DDK k = make_ddk(p);
    k.dead_cap = dead_cap;
    dd_reinject_philox_log_k<<<64, 256, 0, (cudaStream_t)stream>>>(k, dead_buf, x0, u0, v0, w0, active, sigma[0], sigma[1], seed,
                                                                  step, global_offset, orig);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_dd_step_prologue(const pic_dd_params* p, const pic_dd_prologue* a, void* stream) {
    PIC_REQUIRE(p && a, "dd_step_prologue: null pointer");
    PIC_REQUIRE(a->Es && a->E0 && a->wall_cum && a->stats && a->nstats > 0 && a->ctl, "dd_step_prologue: step state missing");
    PIC_REQUIRE(a->n_draws >= 0, "dd_step_prologue: n_draws<0");
    PIC_REQUIRE(!(a->philox && a->n_draws), "dd_step_prologue: device and host draws are exclusive");
    PIC_REQUIRE(a->next_log != a->log || !a->philox, "dd_step_prologue: the log read and the log reset must differ");
    if (a->philox) PIC_REQUIRE(a->log && a->log_cap > 0 && a->x0 && a->u0 && a->active, "dd_step_prologue: philox re-injection arguments");
    if (a->n_draws) {
        PIC_REQUIRE(a->slot && a->ud && a->u0 && a->active, "dd_step_prologue: host draws missing");
        PIC_REQUIRE((!a->xd || a->x0) && (!a->v0 || a->vd) && (!a->w0 || a->wd), "dd_step_prologue: draws / targets missing");
    }
    DDK k = make_ddk(p);
    k.dead_cap = a->log_cap;
    DDPro d;
    d.log = a->log; d.next_log = a->next_log; d.philox = a->philox ? 1 : 0;
    d.s0 = a->sigma[0]; d.s1 = a->sigma[1]; d.seed = a->seed; d.step = a->step; d.goff = a->global_offset;
    d.slot = a->slot; d.dorig = a->orig_of_draw; d.xd = a->xd; d.ud = a->ud; d.vd = a->vd; d.wd = a->wd;
    d.n_draws = a->n_draws; d.corr = a->corr;
    d.x0 = a->x0; d.u0 = a->u0; d.v0 = a->v0; d.w0 = a->w0; d.active = a->active; d.oid = a->orig;
    d.Es = a->Es; d.E0 = a->E0; d.wall_cum = a->wall_cum; d.stats = a->stats; d.nstats = a->nstats; d.ctl = a->ctl;
    int blocks = (k.Ng + 255) / 256;
    blocks = blocks < 16 ? 16 : (blocks > 592 ? 592 : blocks);
    dd_step_prologue_k<<<blocks, 256, 0, (cudaStream_t)stream>>>(k, d);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_dd_thermostat_philox(const pic_dd_params* p, double* u0, double* v0, double* w0, const int8_t* active,
                                 const int32_t* orig, double gamma, const double sigma[2], uint64_t seed, uint64_t step,
                                 int64_t global_offset, void* stream) {
    PIC_REQUIRE(p && u0 && active && sigma, "dd_thermostat_philox: null pointer");
    if (p->N == 0 || !(gamma > 0.0)) return PIC_OK;
    dd_thermostat_philox_k<<<grid_for(p->N, 256, 8), 256, 0, (cudaStream_t)stream>>>(make_ddk(p), u0, v0, w0, active, orig, gamma,
                                                                                    sigma[0], sigma[1], seed, step, global_offset);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gather_i32(const int32_t* src, const int32_t* idx, int32_t* dst, int64_t n, void* stream) {
    PIC_REQUIRE(n >= 0, "gather_i32: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(src && idx && dst, "gather_i32: null pointer");
    gather_i32_k<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_scatter_f64(const double* src, const int32_t* idx, double* dst, int64_t n, void* stream) {
    PIC_REQUIRE(n >= 0, "scatter_f64: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(src && dst, "scatter_f64: null pointer");
    scatter_k<double><<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_scatter_i8(const int8_t* src, const int32_t* idx, int8_t* dst, int64_t n, void* stream) {
    PIC_REQUIRE(n >= 0, "scatter_i8: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(src && dst, "scatter_i8: null pointer");
    scatter_k<int8_t><<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_invert_perm(const int32_t* perm, int32_t* inv, int64_t n, void* stream) {
    PIC_REQUIRE(n >= 0, "invert_perm: n<0");
    if (n == 0) return PIC_OK;
    PIC_REQUIRE(perm && inv, "invert_perm: null pointer");
    invert_perm_k<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(perm, inv, n);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_moments(const double* u, int64_t N, double* out2, void* stream) {
    PIC_REQUIRE(u && out2 && N >= 0, "moments: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    PIC_CHECK_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(double), st));
    if (N == 0) return PIC_OK;
    int grid = grid_for(N, 256, 8);
    if (grid > MOM_MAX_CTAS) grid = MOM_MAX_CTAS;
    moments_k<<<grid, 256, 0, st>>>(u, N, out2);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

}  // extern "C"
