// Device-side initialisers and the IEAD wall-hit histogram (N1).
//
// Initialisers draw from the counter-based Philox4x32-10 generator keyed by (seed, stream) and
// the GLOBAL particle index, so the state of a run does not depend on how the particles are
// sharded over ranks.  They reproduce the reference's DISTRIBUTIONS (x ~ U, v ~ N(mean, sigma),
// the cosine-weighted perturbation loader of pypic.initialize_p), not its MT19937 stream -- the
// host initialisers (pypic.initialize_p, PIC_L_DD.initialize, Particle._initialize_6D in the
// drop-in modules) keep draw-order parity for seeded parity runs.
#include "common.cuh"
#include "host_common.h"

namespace pic {

__device__ __forceinline__ void box_muller(double u1, double u2, double& z0, double& z1) {
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = r * c; z1 = r * s;
}

// x ~ U(xlo, xhi); up to three velocity components ~ N(mean_s, sigma_s) with s = species of the slot
// (PIC_L_DD.initialize 'beam' :223-314, Particle._initialize_6D pygcpic.py:277-304)
__global__ void init_uniform_maxwellian_k(double* __restrict__ x, double* __restrict__ v0, double* __restrict__ v1,
                                          double* __restrict__ v2, long long N, long long n_split, double xlo, double xhi,
                                          double s0, double s1, double m0, double m1, uint64_t seed, uint64_t stream_id,
                                          long long goff) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const uint64_t g = (uint64_t)(goff + i);
        uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x5bd1e995u);
        uint32_t d[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)stream_id ^ 0x9e3779b9u, (uint32_t)(stream_id >> 32)};
        philox4x32(d, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x5bd1e995u);
        const bool sp = i >= n_split;
        const double sig = sp ? s1 : s0, mean = sp ? m1 : m0;
        if (x) x[i] = xlo + u53(c[0], c[1]) * (xhi - xlo);
        double z0, z1, z2, z3;
        box_muller(u53(c[2], c[3]), u53(d[0], d[1]), z0, z1);
        if (v0) v0[i] = mean + sig * z0;
        if (v1) v1[i] = sig * z1;
        if (v2) { box_muller(u53(d[2], d[3]), u53(c[1], d[2]), z2, z3); v2[i] = sig * z2; }
    }
}

// pypic.initialize_p :457-467: the first sum(int(F[i])) particles are placed uniformly inside cell
// i, int(F[i]) of them per cell in cell order; prefix[i] = sum_{j<i} int(F[j]) (Ng+1 entries).
__global__ void pypic_perturb_positions_k(double* __restrict__ x, long long N, const long long* __restrict__ prefix,
                                          const double* __restrict__ X, int Ng, uint64_t seed, long long goff) {
    const long long total = prefix[Ng];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const long long g = goff + i;
        if (g >= total) continue;
        int lo = 0, hi = Ng;                               // largest cell with prefix[cell] <= g
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= g) lo = mid; else hi = mid; }
        uint32_t c[4] = {(uint32_t)g, (uint32_t)((uint64_t)g >> 32), 0x70657274u, 0u};
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x5bd1e995u);
        x[i] = X[lo] + u53(c[0], c[1]) * (X[lo + 1] - X[lo]);
    }
}

// IEAD (pygcpic.py:1574-1584): 2-D histogram of (kinetic_energy/e, angle w.r.t. the wall normal) of
// the particles flagged in `sel` (e.g. hit_flag) whose Z matches, with numpy.histogram2d semantics
// (bin b holds edges[b] <= v < edges[b+1], the last bin is closed on the right).
__device__ __forceinline__ int hist_bin(const double* __restrict__ e, int nb, double v) {
    if (!(v >= e[0]) || v > e[nb]) return -1;
    if (v == e[nb]) return nb - 1;
    int lo = 0, hi = nb;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (e[mid] <= v) lo = mid; else hi = mid; }
    return lo;
}
__global__ void gc_iead_hist_k(const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                               const double* __restrict__ m, const int8_t* __restrict__ sel, const int32_t* __restrict__ Z,
                               int Zsel, long long N, const double* __restrict__ e_edges, int ne,
                               const double* __restrict__ a_edges, int na, double* __restrict__ hist) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        if (sel[i] != 1 || (Z && Z[i] != Zsel)) continue;
        const double a = vx[i], b = vy[i], c = vz[i];
        const double speed = sqrt(a * a + b * b + c * c);               // pygcpic.py:213-226
        const double ke = 0.5 * m[i] * (speed * speed) / PIC_E;         // :262-275, /e at :1518
        const double ang = atan2(sqrt(b * b + c * c), fabs(a)) * 180. / 3.141592653589793;   // :228-259
        const int be = hist_bin(e_edges, ne, ke), ba = hist_bin(a_edges, na, ang);
        if (be >= 0 && ba >= 0) atomicAdd(&hist[(long long)be * na + ba], 1.0);
    }
}

}  // namespace pic

using namespace pic;

extern "C" {

int pic_dev_init_uniform_maxwellian(double* x, double* v0, double* v1, double* v2, int64_t N, int64_t n_split, double xlo,
                                    double xhi, const double sigma[2], const double mean[2], uint64_t seed,
                                    uint64_t stream_id, int64_t global_offset, void* stream) {
    PIC_REQUIRE(N >= 0 && sigma && mean && xhi >= xlo, "init_uniform_maxwellian: bad argument");
    if (N == 0) return PIC_OK;
    init_uniform_maxwellian_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, v0, v1, v2, N, n_split, xlo, xhi, sigma[0],
                                                                                     sigma[1], mean[0], mean[1], seed, stream_id,
                                                                                     global_offset);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_pypic_perturb_positions(double* x, int64_t N, const int64_t* prefix, const double* X, int Ng, uint64_t seed,
                                    int64_t global_offset, void* stream) {
    PIC_REQUIRE(x && prefix && X && N >= 0 && Ng >= 1, "pypic_perturb_positions: bad argument");
    if (N == 0) return PIC_OK;
    pypic_perturb_positions_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, N, (const long long*)prefix, X, Ng, seed,
                                                                                     global_offset);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

int pic_dev_gc_iead_hist(const double* vx, const double* vy, const double* vz, const double* m, const int8_t* select,
                         const int32_t* Z, int Z_select, int64_t N, const double* e_edges, int n_e_bins,
                         const double* a_edges, int n_a_bins, double* hist, void* stream) {
    PIC_REQUIRE(vx && vy && vz && m && select && e_edges && a_edges && hist && N >= 0 && n_e_bins >= 1 && n_a_bins >= 1,
                "gc_iead_hist: bad argument");
    if (N == 0) return PIC_OK;
    gc_iead_hist_k<<<grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(vx, vy, vz, m, select, Z, Z_select, N, e_edges, n_e_bins,
                                                                          a_edges, n_a_bins, hist);
    PIC_CHECK_LAUNCH();
    return PIC_OK;
}

}  // extern "C"
