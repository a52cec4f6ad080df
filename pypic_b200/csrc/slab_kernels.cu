// Field phase of the bounded sheath on SPATIAL SLABS (pypic_b200/spatial.py; BASELINE config 5): every rank
// updates the field on its own nodes plus `G` guard nodes either side, so what travels per Picard iteration is
// two guard bands, two partial sums and four counts per rank instead of the whole grid.
//
// PIC_L_DD.py:55-66 (wall terms, edge fold) and :516-527 (E1 = E0 + dt/eps0 (mean(jh) - jh), Eh, residual) are
// the formulas of dd_field_update_k (dd_kernels.cu); what differs is where the operands come from:
//   * rank r's particles live in the cells [c0, c1) (+- G cells of drift between two sorts), so its RAW
//     accumulators are non-zero only on the band [c0-G, c1+G];
//   * the complete current at a node within G of a slab boundary is the sum of the two neighbours' raw values
//     there (slabs are wider than 2G+2 cells, so no third rank reaches it).  Each rank sends the raw values of
//     the 2G+1 nodes around each of its boundaries; both neighbours then form the same sum (one commutative
//     addition) and update E there redundantly, with identical bits -- no second exchange for the guard nodes
//     of E;
//   * mean(jh) needs the global sum: sum_i jh_i = sum_r S_r + wallL + wallR + jh[1] + jh[Ng-2], where S_r is
//     the sum of ALL raw deposits of rank r (formed before the exchange) and the last two terms are the fold
//     j[0] += j[1], j[-1] += j[-2], known to ranks 0 and W-1, who add them to their S;
//   * the residual and the field energy are sums over OWNED nodes, completed by a second exchange of two doubles.
// All partial sums are formed in a fixed order (per-CTA partials summed by the last CTA in CTA order, ranks in
// rank order), so every rank computes the same mean and the same residual, and takes the same Picard exit.
//
// Message of rank r (fp64[4B+10], B = 2G+1):
//   [jh on c0-G..c0+G | j1 there | jh on c1-G..c1+G | j1 there | S_h | S_1 | 4 absorbed counts | jh[1], j1[1] (rank 0) |
//    jh[Ng-2], j1[Ng-2] (rank W-1)]; the bands of the outer ranks towards the walls are zero.
#include "common.cuh"
#include "host_common.h"

namespace pic {

#define SLAB_MAXP 128           // CTAs of the pack / field kernels (per-CTA partial sums)

struct SlabK {
    int Ng, c0, c1, G, rank, world;
    int b0, b1;                 // band [b0, b1): the nodes this rank's particles may touch
    int o0, o1;                 // owned nodes [o0, o1) (the last rank also owns node Ng-1)
    double dx, dt, p2c, q[2];
};

__device__ __forceinline__ int slab_msg_len(int G) { return 4 * (2 * G + 1) + 10; }

// last-CTA-done: returns true in every thread of the CTA that arrives last
__device__ __forceinline__ bool slab_last_cta(unsigned* ticket, int* s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        *s_flag = (t == gridDim.x - 1);
        if (*s_flag) *ticket = 0;            // ready for the next launch (stream-ordered)
        __threadfence();
    }
    __syncthreads();
    return *s_flag != 0;
}

__global__ void __launch_bounds__(1024) slab_pack_k(SlabK k, double* __restrict__ acc, double* __restrict__ msg,
                                                    double* __restrict__ work, double* __restrict__ absorbed_local,
                                                    const int* __restrict__ ctl) {
    __shared__ double scratch[33];
    __shared__ int s_last;
    if (ctl && *(volatile const int*)ctl) return;
    const int Ng = k.Ng, B = 2 * k.G + 1;
    const double* jh = acc;
    const double* j1 = acc + Ng;
    double sh = 0.0, s1 = 0.0;
    for (int i = k.b0 + blockIdx.x * blockDim.x + threadIdx.x; i < k.b1; i += gridDim.x * blockDim.x) { sh += jh[i]; s1 += j1[i]; }
    sh = block_reduce<0>(sh, scratch);
    s1 = block_reduce<0>(s1, scratch);
    if (threadIdx.x == 0) { work[2 * blockIdx.x] = sh; work[2 * blockIdx.x + 1] = s1; }
    if (blockIdx.x == 0) {
        // the two boundary bands (zero towards a wall) and the fold nodes
        for (int t = threadIdx.x; t < 2 * B; t += blockDim.x) {
            const int a = t >= B, n = t - a * B;
            const bool haveL = k.rank > 0, haveR = k.rank < k.world - 1;
            const int iL = k.c0 - k.G + n, iR = k.c1 - k.G + n;
            msg[a * B + n] = haveL ? acc[a * Ng + iL] : 0.0;
            msg[2 * B + a * B + n] = haveR ? acc[a * Ng + iR] : 0.0;
        }
        if (threadIdx.x < 4) msg[4 * B + 6 + threadIdx.x] = 0.0;
        __syncthreads();
        if (threadIdx.x < 2) {
            const int a = threadIdx.x;
            if (k.rank == 0) msg[4 * B + 6 + a] = acc[a * Ng + 1];
            if (k.rank == k.world - 1) msg[4 * B + 8 + a] = acc[a * Ng + Ng - 2];
        }
    }
    unsigned* ticket = (unsigned*)(work + 2 * SLAB_MAXP);
    if (slab_last_cta(ticket, &s_last)) {
        if (threadIdx.x < 2) {
            const int a = threadIdx.x;
            double s = 0.0;
            for (int c = 0; c < (int)gridDim.x; ++c) s += __ldcg(&work[2 * c + a]);
            if (k.rank == 0) s += acc[a * Ng + 1];
            if (k.rank == k.world - 1) s += acc[a * Ng + Ng - 2];
            msg[4 * B + a] = s;
        }
        // absorbed counts of this iteration; cleared here (the field kernel clears the band)
        if (threadIdx.x >= 32 && threadIdx.x < 36) {
            const int c = threadIdx.x - 32;
            const double n = acc[2 * Ng + c];
            msg[4 * B + 2 + c] = n;
            if (absorbed_local) absorbed_local[c] += n;      // what THIS rank absorbed in the step (re-injection)
            acc[2 * Ng + c] = 0.0;
        }
    }
}

// gath = the messages of all ranks in rank order.  send2[0..1] = this rank's partial (residual^2, field energy).
__global__ void __launch_bounds__(1024) slab_field_k(SlabK k, double* __restrict__ acc, const double* __restrict__ gath,
                                                     const double* __restrict__ wall_cum, const double* __restrict__ E0,
                                                     double* __restrict__ Es, double* __restrict__ E1,
                                                     double* __restrict__ j1o, double* __restrict__ send2,
                                                     double* __restrict__ work, const int* __restrict__ ctl) {
    __shared__ double scratch[33];
    __shared__ double s_c[3];
    __shared__ int s_last;
    if (ctl && *(volatile const int*)ctl) return;
    const int Ng = k.Ng, B = 2 * k.G + 1, M = 4 * B + 10;
    if (threadIdx.x == 0) {
        double sh = 0.0, w[4] = {wall_cum[0], wall_cum[1], wall_cum[2], wall_cum[3]};
        for (int r = 0; r < k.world; ++r) {
            const double* m = gath + (size_t)r * M + 4 * B;
            sh += m[0];
            for (int c = 0; c < 4; ++c) w[c] += m[2 + c];
        }
        // wall-charge current of every particle absorbed so far in this step (PIC_L_DD.py:58,62): count * value
        const double wallL = w[0] * (k.dx * k.q[0] * k.p2c / k.dt) + w[1] * (k.dx * k.q[1] * k.p2c / k.dt);
        const double wallR = w[2] * (-k.dx * k.q[0] * k.p2c / k.dt) + w[3] * (-k.dx * k.q[1] * k.p2c / k.dt);
        s_c[0] = ((sh + wallL) + wallR) / (double)Ng;
        s_c[1] = wallL; s_c[2] = wallR;
    }
    __syncthreads();
    const double meanh = s_c[0], wallL = s_c[1], wallR = s_c[2];
    const double coef = k.dt / PIC_EPS0;
    const double* left = k.rank > 0 ? gath + (size_t)(k.rank - 1) * M + 2 * B : nullptr;           // r-1's band around c0
    const double* right = k.rank < k.world - 1 ? gath + (size_t)(k.rank + 1) * M : nullptr;        // r+1's band around c1
    const double* mine = gath + (size_t)k.rank * M + 4 * B + 6;                                    // fold nodes
    double rr = 0.0, ee = 0.0;
    for (int i = k.b0 + blockIdx.x * blockDim.x + threadIdx.x; i < k.b1; i += gridDim.x * blockDim.x) {
        double a = acc[i], b = acc[Ng + i];
        acc[i] = 0.0; acc[Ng + i] = 0.0;
        if (left) { const unsigned n = (unsigned)(i - (k.c0 - k.G)); if (n < (unsigned)B) { a += left[n]; b += left[B + n]; } }
        if (right) { const unsigned n = (unsigned)(i - (k.c1 - k.G)); if (n < (unsigned)B) { a += right[n]; b += right[B + n]; } }
        // edge fold j[0]+=j[1]; j[-1]+=j[-2] with the unfolded neighbours (:65-66)
        if (i == 0) { a = (a + wallL) + mine[0]; b = (b + wallL) + mine[1]; }
        if (i == Ng - 1) { a = (a + wallR) + mine[2]; b = (b + wallR) + mine[3]; }
        const double e0 = E0[i];
        const double e1 = e0 + coef * (meanh - a);      // :516
        const double eh = (e1 + e0) * 0.5;               // :521
        const double d = Es[i] - eh;
        if (i >= k.o0 && i < k.o1) { rr += d * d; ee += PIC_EPS0 * e1 * e1 * k.dx / 2.; }
        E1[i] = e1;
        Es[i] = eh;
        j1o[i] = b;
    }
    rr = block_reduce<0>(rr, scratch);
    ee = block_reduce<0>(ee, scratch);
    double* part = work + 2 * SLAB_MAXP + 2;
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = rr; part[2 * blockIdx.x + 1] = ee; }
    unsigned* ticket = (unsigned*)(work + 2 * SLAB_MAXP) + 1;
    if (slab_last_cta(ticket, &s_last) && threadIdx.x < 2) {
        double s = 0.0;
        for (int c = 0; c < (int)gridDim.x; ++c) s += __ldcg(&part[2 * c + threadIdx.x]);
        send2[threadIdx.x] = s;
    }
}

// gath2 = [residual^2, field energy] of every rank.  Closes the iteration: counts, statistics, loop flag.
__global__ void slab_finish_k(SlabK k, const double* __restrict__ gath, const double* __restrict__ gath2,
                              double* __restrict__ wall_cum, double* __restrict__ stats, double* __restrict__ rhist,
                              int* __restrict__ ctl, double tol, int maxiter) {
    if (ctl && *(volatile int*)ctl) return;
    if (threadIdx.x != 0) return;
    const int B = 2 * k.G + 1, M = 4 * B + 10;
    double rr = 0.0, ee = 0.0, s1 = 0.0, w[4] = {wall_cum[0], wall_cum[1], wall_cum[2], wall_cum[3]};
    for (int r = 0; r < k.world; ++r) {
        rr += gath2[2 * r]; ee += gath2[2 * r + 1];
        const double* m = gath + (size_t)r * M + 4 * B;
        s1 += m[1];
        for (int c = 0; c < 4; ++c) w[c] += m[2 + c];
    }
    for (int c = 0; c < 4; ++c) wall_cum[c] = w[c];
    const double wallL = w[0] * (k.dx * k.q[0] * k.p2c / k.dt) + w[1] * (k.dx * k.q[1] * k.p2c / k.dt);
    const double wallR = w[2] * (-k.dx * k.q[0] * k.p2c / k.dt) + w[3] * (-k.dx * k.q[1] * k.p2c / k.dt);
    const double r = sqrt(rr);                     // np.linalg.norm(Es-Eh), :525
    const double it = stats[3] + 1.0;
    stats[0] = r;
    stats[1] = ((s1 + wallL) + wallR) / (double)k.Ng;      // np.average(j1) -> jbias, :551
    stats[2] = ee;                                 // sum(eps0*E*E*dx/2), :548
    stats[3] = it;
    if (rhist && it <= (double)maxiter) rhist[(int)it - 1] = r;
    if (ctl && (!(r > tol) || it >= (double)maxiter)) *ctl = 1;      // `while r > tol and k < maxiter`, :452
}

static int make_slabk(const pic_dd_params* p, int c0, int c1, int guard, int rank, int world, SlabK* k) {
    PIC_REQUIRE(p && p->Ng >= 3 && p->dx > 0 && p->dt > 0, "slab field: bad parameters");
    PIC_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && guard >= 1, "slab field: bad rank / world / guard");
    PIC_REQUIRE(c0 >= 0 && c1 > c0 && c1 <= p->Ng - 1, "slab field: bad slab bounds");
    PIC_REQUIRE((rank == 0) == (c0 == 0) && (rank == world - 1) == (c1 == p->Ng - 1), "slab field: outer slabs must touch the walls");
    PIC_REQUIRE(world == 1 || c1 - c0 > 2 * guard + 2, "slab field: slabs must be wider than 2*guard+2 cells");
    k->Ng = p->Ng; k->c0 = c0; k->c1 = c1; k->G = guard; k->rank = rank; k->world = world;
    k->b0 = c0 - guard < 0 ? 0 : c0 - guard;
    k->b1 = c1 + guard + 1 > p->Ng ? p->Ng : c1 + guard + 1;
    k->o0 = c0; k->o1 = rank == world - 1 ? c1 + 1 : c1;
    k->dx = p->dx; k->dt = p->dt; k->p2c = p->p2c; k->q[0] = p->q[0]; k->q[1] = p->q[1];
    return PIC_OK;
}

static int slab_grid(const SlabK& k) {
    int g = (k.b1 - k.b0 + 4095) / 4096;
    if (g > SLAB_MAXP) g = SLAB_MAXP;
    if (g > device_sm_count()) g = device_sm_count();
    return g < 1 ? 1 : g;
}

}  // namespace pic

using namespace pic;

extern "C" {

int pic_slab_message_len(int guard) { return 4 * (2 * guard + 1) + 10; }
int pic_slab_work_len(void) { return 4 * SLAB_MAXP + 2; }

int pic_dev_slab_pack(const pic_dd_params* p, int c0, int c1, int guard, int rank, int world, double* acc, double* msg,
                      double* work, double* absorbed_local, const int32_t* ctl, void* stream) {
    PIC_REQUIRE(acc && msg && work, "slab_pack: null pointer");
    SlabK k;
    int rc = make_slabk(p, c0, c1, guard, rank, world, &k);
    if (rc) return rc;
    slab_pack_k<<<slab_grid(k), 1024, 0, (cudaStream_t)stream>>>(k, acc, msg, work, absorbed_local, ctl);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_slab_field_update(const pic_dd_params* p, int c0, int c1, int guard, int rank, int world, double* acc,
                              const double* gathered, const double* wall_cum, const double* E0, double* Es, double* E1,
                              double* j1, double* partial, double* work, const int32_t* ctl, void* stream) {
    PIC_REQUIRE(acc && gathered && wall_cum && E0 && Es && E1 && j1 && partial && work, "slab_field_update: null pointer");
    SlabK k;
    int rc = make_slabk(p, c0, c1, guard, rank, world, &k);
    if (rc) return rc;
    slab_field_k<<<slab_grid(k), 1024, 0, (cudaStream_t)stream>>>(k, acc, gathered, wall_cum, E0, Es, E1, j1, partial, work, ctl);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_slab_finish(const pic_dd_params* p, int c0, int c1, int guard, int rank, int world, const double* gathered,
                        const double* partials, double* wall_cum, double* stats, double* rhist, int32_t* ctl, double tol,
                        int maxiter, void* stream) {
    PIC_REQUIRE(gathered && partials && wall_cum && stats, "slab_finish: null pointer");
    PIC_REQUIRE(!(ctl || rhist) || maxiter >= 1, "slab_finish: maxiter must be >= 1 with ctl / rhist");
    SlabK k;
    int rc = make_slabk(p, c0, c1, guard, rank, world, &k);
    if (rc) return rc;
    slab_finish_k<<<1, 32, 0, (cudaStream_t)stream>>>(k, gathered, partials, wall_cum, stats, rhist, ctl, tol, maxiter);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

}  // extern "C"
