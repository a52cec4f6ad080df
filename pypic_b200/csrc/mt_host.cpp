// Host draw service for the RNG-parity mode of the sheath (PIC_L_DD.py:419-450).
//
// The reference draws from NumPy's legacy global stream (MT19937 + legacy_gauss) in particle
// index order: one uniform per ACTIVE particle for the thermostat (even when gamma == 0: Python
// evaluates `active[i]==1 and np.random.uniform(0,1) < gamma` left to right), then
// x = uniform(0,L) and u,v,w = normal(0,sigma) per dead slot.  A device-resident run needs two
// things from that stream without paying 2 words per particle per step:
//   * SKIP the thermostat uniforms: MT19937 is linear over GF(2), so advancing the state by J
//     words is the polynomial g(t) = t^J mod phi(t) applied to the state (phi = characteristic
//     polynomial of the recurrence, degree 19937).  In sequence form: z[n+J] = XOR_{i: g_i=1} z[n+i]
//     for the raw (untempered) word sequence z, so the new 624-word state is a GF(2) convolution
//     of 19937+624 words generated from the old one -- ~1 ms instead of ~60 ms at 2e7 particles;
//   * the few hundred re-injection draws per step, bit-identical to np.random.uniform / normal
//     (legacy polar Box-Muller with its cached second variate, libm log/sqrt like NumPy's C core).
// phi is obtained once by Berlekamp-Massey on one output bit; g by square-and-multiply.
// Nothing here touches the GPU; it is compiled into libpic_b200.so with the host compiler.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <mutex>
#include <vector>

#include "../../include/pic_b200.h"
#include "host_common.h"

namespace {

const int MT_N = 624, MT_M = 397, MT_DEG = 19937;
const uint32_t MT_A = 0x9908b0dfu, MT_UP = 0x80000000u, MT_LOW = 0x7fffffffu;
const int PW = 312;                      // 64-bit words of a polynomial of degree < 19968

inline uint32_t mt_twist(uint32_t a, uint32_t b) {
    const uint32_t y = (a & MT_UP) | (b & MT_LOW);
    return (y >> 1) ^ ((y & 1u) ? MT_A : 0u);
}

// NumPy's legacy generator (numpy/random/src/mt19937/mt19937.c: mt19937_gen + tempering)
struct MT {
    uint32_t* key;
    int pos;
    void regen() {
        int i;
        for (i = 0; i < MT_N - MT_M; ++i) key[i] = key[i + MT_M] ^ mt_twist(key[i], key[i + 1]);
        for (; i < MT_N - 1; ++i) key[i] = key[i + (MT_M - MT_N)] ^ mt_twist(key[i], key[i + 1]);
        key[MT_N - 1] = key[MT_M - 1] ^ mt_twist(key[MT_N - 1], key[0]);
        pos = 0;
    }
    inline uint32_t next32() {
        if (pos == MT_N) regen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    inline double next_double() {                 // mt19937_next_double
        const int32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

struct Gauss {                                     // legacy_gauss (legacy-distributions.c)
    int has;
    double val;
    inline double next(MT& g) {
        if (has) { const double t = val; has = 0; val = 0.0; return t; }
        double f, x1, x2, r2;
        do {
            x1 = 2.0 * g.next_double() - 1.0;
            x2 = 2.0 * g.next_double() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        f = sqrt(-2.0 * log(r2) / r2);
        val = f * x1; has = 1;
        return f * x2;
    }
};

// ---- GF(2) polynomial helpers (bit i of word i/64 = coefficient of t^i) ----
inline int getbit(const uint64_t* p, int i) { return (int)((p[i >> 6] >> (i & 63)) & 1u); }
inline void xor_shifted(uint64_t* dst, const uint64_t* src, int nsrc_words, int shift) {
    const int q = shift >> 6, r = shift & 63;
    if (r == 0) { for (int w = 0; w < nsrc_words; ++w) dst[q + w] ^= src[w]; return; }
    uint64_t carry = 0;
    for (int w = 0; w < nsrc_words; ++w) {
        dst[q + w] ^= (src[w] << r) | carry;
        carry = src[w] >> (64 - r);
    }
    dst[q + nsrc_words] ^= carry;
}
inline uint64_t extract64(const uint64_t* p, long bit) {        // 64 bits starting at `bit` (p padded by one word)
    const long q = bit >> 6; const int r = (int)(bit & 63);
    return r ? (p[q] >> r) | (p[q + 1] << (64 - r)) : p[q];
}
inline uint64_t spread32(uint32_t v) {                          // bit i -> bit 2i
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// y[m] = XOR over the set bits i of g of Z[i+m], m < 624: the GF(2) convolution that applies the
// jump polynomial to the raw word sequence (~10^4 set bits x 624 words).  The output is produced in
// blocks of six vector registers that stay in registers while the set bits are walked, so the loop
// is nothing but unaligned loads and XORs: 0.33 ms with AVX2, 0.55 ms with SSE2 (chosen at run time).
#if defined(__x86_64__)
__attribute__((target("avx2")))
void jump_convolve_avx2(const int* idx, int n, const uint32_t* Z, uint32_t* y) {
    for (int mb = 0; mb < MT_N; mb += 48) {             // 624 = 13 x 48
        __m256i a0 = _mm256_setzero_si256(), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0;
        for (int k = 0; k < n; ++k) {
            const __m256i* z = (const __m256i*)(Z + idx[k] + mb);
            a0 = _mm256_xor_si256(a0, _mm256_loadu_si256(z));     a1 = _mm256_xor_si256(a1, _mm256_loadu_si256(z + 1));
            a2 = _mm256_xor_si256(a2, _mm256_loadu_si256(z + 2)); a3 = _mm256_xor_si256(a3, _mm256_loadu_si256(z + 3));
            a4 = _mm256_xor_si256(a4, _mm256_loadu_si256(z + 4)); a5 = _mm256_xor_si256(a5, _mm256_loadu_si256(z + 5));
        }
        __m256i* o = (__m256i*)(y + mb);
        _mm256_storeu_si256(o, a0); _mm256_storeu_si256(o + 1, a1); _mm256_storeu_si256(o + 2, a2);
        _mm256_storeu_si256(o + 3, a3); _mm256_storeu_si256(o + 4, a4); _mm256_storeu_si256(o + 5, a5);
    }
}
void jump_convolve_sse2(const int* idx, int n, const uint32_t* Z, uint32_t* y) {
    for (int mb = 0; mb < MT_N; mb += 24) {             // 624 = 26 x 24
        __m128i a0 = _mm_setzero_si128(), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0;
        for (int k = 0; k < n; ++k) {
            const __m128i* z = (const __m128i*)(Z + idx[k] + mb);
            a0 = _mm_xor_si128(a0, _mm_loadu_si128(z));     a1 = _mm_xor_si128(a1, _mm_loadu_si128(z + 1));
            a2 = _mm_xor_si128(a2, _mm_loadu_si128(z + 2)); a3 = _mm_xor_si128(a3, _mm_loadu_si128(z + 3));
            a4 = _mm_xor_si128(a4, _mm_loadu_si128(z + 4)); a5 = _mm_xor_si128(a5, _mm_loadu_si128(z + 5));
        }
        __m128i* o = (__m128i*)(y + mb);
        _mm_storeu_si128(o, a0); _mm_storeu_si128(o + 1, a1); _mm_storeu_si128(o + 2, a2);
        _mm_storeu_si128(o + 3, a3); _mm_storeu_si128(o + 4, a4); _mm_storeu_si128(o + 5, a5);
    }
}
#endif
void jump_convolve(const int* idx, int n, const uint32_t* Z, uint32_t* y) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) jump_convolve_avx2(idx, n, Z, y);
    else jump_convolve_sse2(idx, n, Z, y);
#else
    memset(y, 0, MT_N * sizeof(uint32_t));
    for (int k = 0; k < n; ++k) { const uint32_t* zi = Z + idx[k]; for (int m = 0; m < MT_N; ++m) y[m] ^= zi[m]; }
#endif
}

std::mutex g_mu;
std::vector<uint64_t> g_phi;        // characteristic polynomial, PW+1 words (degree 19937), empty until computed

// Berlekamp-Massey on bit 0 of the raw word sequence of an arbitrary non-zero state
int compute_phi() {
    if (!g_phi.empty()) return PIC_OK;
    const int LEN = 2 * MT_DEG + 64;
    std::vector<uint32_t> z(MT_N + LEN);
    uint32_t s = 19650218u;                                  // init_genrand
    for (int i = 0; i < MT_N; ++i) { z[i] = s; s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(i + 1); }
    for (int k = 0; k < LEN; ++k) z[MT_N + k] = z[MT_M + k] ^ mt_twist(z[k], z[k + 1]);
    // sequence s[n] = bit 0 of z[624+n]; stored reversed: srev bit (LEN-1-n) = s[n]
    const int SW = (LEN + 63) / 64 + 2;
    std::vector<uint64_t> srev(SW, 0);
    for (int n = 0; n < LEN; ++n)
        if (z[MT_N + n] & 1u) { const int p = LEN - 1 - n; srev[p >> 6] |= 1ull << (p & 63); }
    const int CW = (LEN + 63) / 64 + 2;
    std::vector<uint64_t> Cp(CW, 0), Bp(CW, 0), Tp(CW, 0);
    Cp[0] = 1; Bp[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < LEN; ++n) {
        // d = sum_{i=0..L} C_i s[n-i];  s[n-i] sits at srev bit (LEN-1-n)+i
        const long o = LEN - 1 - n;
        uint64_t acc = 0;
        const int nw = (L >> 6) + 1;
        for (int w = 0; w < nw; ++w) acc ^= Cp[w] & extract64(srev.data(), o + 64l * w);
        const int d = __builtin_parityll(acc);
        if (!d) { ++m; continue; }
        const int bw = ((n - m >= 0 ? LEN : LEN) >> 6) + 1;   // words of B in use (bounded by its degree <= n)
        (void)bw;
        const int bwords = (n >> 6) + 2 < CW - 1 - ((m + 63) >> 6) ? (n >> 6) + 2 : CW - 1 - ((m + 63) >> 6);
        if (2 * L <= n) {
            Tp = Cp;
            xor_shifted(Cp.data(), Bp.data(), bwords, m);
            L = n + 1 - L;
            Bp.swap(Tp);
            m = 1;
        } else {
            xor_shifted(Cp.data(), Bp.data(), bwords, m);
            ++m;
        }
    }
    if (L != MT_DEG) { pic::set_error("mt: Berlekamp-Massey found linear complexity %d, expected 19937", L); return PIC_ERR_ARG; }
    // connection polynomial C(x) -> characteristic polynomial phi(t) = t^L C(1/t): phi_j = c_{L-j}
    std::vector<uint64_t> phi(PW + 1, 0);
    for (int j = 0; j <= MT_DEG; ++j)
        if (getbit(Cp.data(), MT_DEG - j)) phi[j >> 6] |= 1ull << (j & 63);
    g_phi.swap(phi);
    return PIC_OK;
}

// r (2*PW words, degree < 2*19937) mod phi, in place; the result occupies the low PW words
void reduce_mod_phi(uint64_t* r) {
    const uint64_t* phi = g_phi.data();
    for (int d = 2 * MT_DEG - 2; d >= MT_DEG; --d)
        if (getbit(r, d)) xor_shifted(r, phi, PW, d - MT_DEG);
}

}  // namespace

extern "C" {

int pic_mt_jump_poly(uint64_t nwords, uint32_t* g624) {
    PIC_REQUIRE(g624, "mt_jump_poly: null pointer");
    std::lock_guard<std::mutex> lk(g_mu);
    int rc = compute_phi();
    if (rc) return rc;
    std::vector<uint64_t> r(2 * PW + 2, 0), sq(2 * PW + 2, 0);
    r[0] = 1;
    int top = 63;
    while (top > 0 && !((nwords >> top) & 1ull)) --top;
    for (int b = top; b >= 0; --b) {
        // square
        std::fill(sq.begin(), sq.end(), 0);
        const uint32_t* r32 = (const uint32_t*)r.data();
        for (int w = 0; w < 2 * PW; ++w) sq[w] = spread32(r32[w]);
        reduce_mod_phi(sq.data());
        std::fill(r.begin(), r.end(), 0);
        memcpy(r.data(), sq.data(), PW * sizeof(uint64_t));
        if ((nwords >> b) & 1ull) {
            // multiply by t
            uint64_t carry = 0;
            for (int w = 0; w <= PW; ++w) { const uint64_t nc = r[w] >> 63; r[w] = (r[w] << 1) | carry; carry = nc; }
            if (getbit(r.data(), MT_DEG)) for (int w = 0; w <= PW; ++w) r[w] ^= g_phi[w];
        }
    }
    memcpy(g624, r.data(), PW * sizeof(uint64_t));
    return PIC_OK;
}

int pic_mt_jump(uint32_t* key624, int32_t* pos, const uint32_t* g624) {
    PIC_REQUIRE(key624 && pos && g624, "mt_jump: null pointer");
    PIC_REQUIRE(*pos >= 0 && *pos <= MT_N, "mt_jump: pos outside [0, 624]");
    const int p = *pos;
    const int ZL = MT_DEG + MT_N;                       // words of the raw sequence needed behind position p
    std::vector<uint32_t> B(p + ZL + 64, 0u);          // + padding: the vector loops read whole blocks
    memcpy(B.data(), key624, MT_N * sizeof(uint32_t));
    for (int k = 0; MT_N + k < p + ZL; ++k) B[MT_N + k] = B[MT_M + k] ^ mt_twist(B[k], B[k + 1]);
    const uint32_t* Z = B.data() + p;
    uint32_t y[MT_N];
    const uint64_t* g = (const uint64_t*)g624;
    std::vector<int> idx;
    idx.reserve(MT_DEG);
    for (int w = 0; w < PW; ++w) {
        uint64_t bits = g[w];
        while (bits) { idx.push_back(64 * w + __builtin_ctzll(bits)); bits &= bits - 1; }
    }
    jump_convolve(idx.data(), (int)idx.size(), Z, y);
    memcpy(key624, y, sizeof(y));
    *pos = 0;
    return PIC_OK;
}

int pic_mt_skip(uint32_t* key624, int32_t* pos, uint64_t nwords) {
    PIC_REQUIRE(key624 && pos && *pos >= 0 && *pos <= MT_N, "mt_skip: bad state");
    MT g{key624, *pos};
    while (nwords) {
        if (g.pos == MT_N) g.regen();
        const uint64_t take = nwords < (uint64_t)(MT_N - g.pos) ? nwords : (uint64_t)(MT_N - g.pos);
        g.pos += (int)take; nwords -= take;
    }
    *pos = g.pos;
    return PIC_OK;
}

int pic_mt_sheath_draws(uint32_t* key624, int32_t* pos, int32_t* has_gauss, double* gauss, int64_t n,
                        const double* sigma, double L, double* xd, double* ud, double* vd, double* wd) {
    PIC_REQUIRE(key624 && pos && has_gauss && gauss && n >= 0, "mt_sheath_draws: bad argument");
    PIC_REQUIRE(*pos >= 0 && *pos <= MT_N, "mt_sheath_draws: pos outside [0, 624]");
    PIC_REQUIRE(!xd || (sigma && ud && vd && wd), "mt_sheath_draws: outputs missing");
    MT g{key624, *pos};
    Gauss ga{*has_gauss, *gauss};
    for (int64_t k = 0; k < n; ++k) {
        // np.random.uniform(0.0, L) = 0.0 + (L - 0.0) * next_double; np.random.normal(0.0, s) = 0.0 + s * gauss
        const double ux = 0.0 + (L - 0.0) * g.next_double();
        const double a = ga.next(g), b = ga.next(g), c = ga.next(g);
        if (xd) { const double s = sigma[k]; xd[k] = ux; ud[k] = 0.0 + s * a; vd[k] = 0.0 + s * b; wd[k] = 0.0 + s * c; }
    }
    *pos = g.pos; *has_gauss = ga.has; *gauss = ga.val;
    return PIC_OK;
}

int pic_mt_sheath_thermostat(uint32_t* key624, int32_t* pos, int32_t* has_gauss, double* gauss, int64_t n_active,
                             int64_t k_split, double gamma, double sigma0, double sigma1, int64_t cap,
                             int64_t* hit_k, double* hu, double* hv, double* hw, int64_t* nhits) {
    PIC_REQUIRE(key624 && pos && has_gauss && gauss && nhits && n_active >= 0, "mt_sheath_thermostat: bad argument");
    PIC_REQUIRE(*pos >= 0 && *pos <= MT_N, "mt_sheath_thermostat: pos outside [0, 624]");
    PIC_REQUIRE(cap == 0 || (hit_k && hu && hv && hw), "mt_sheath_thermostat: outputs missing");
    MT g{key624, *pos};
    Gauss ga{*has_gauss, *gauss};
    int64_t nh = 0;
    for (int64_t k = 0; k < n_active; ++k) {
        const double u = 0.0 + (1.0 - 0.0) * g.next_double();
        if (u < gamma) {
            const double s = k < k_split ? sigma0 : sigma1;
            const double a = ga.next(g), b = ga.next(g), c = ga.next(g);
            if (nh < cap) { hit_k[nh] = k; hu[nh] = 0.0 + s * a; hv[nh] = 0.0 + s * b; hw[nh] = 0.0 + s * c; }
            ++nh;
        }
    }
    *nhits = nh;
    if (nh > cap) {      // state is NOT advanced: the caller retries with a larger capacity
        pic::set_error("mt_sheath_thermostat: %lld hits exceed the capacity %lld", (long long)nh, (long long)cap);
        return PIC_ERR_ARG;
    }
    *pos = g.pos; *has_gauss = ga.has; *gauss = ga.val;
    return PIC_OK;
}

}  // extern "C"
