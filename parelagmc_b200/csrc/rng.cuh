// rng.cuh -- GPU stream generator compatible with trng::yarn5 + trng::normal_dist<double>, the pair wrapped by
// NormalDistributionSampler (/root/reference/src/NormalDistributionSampler.hpp:61-64, .cpp:17-37), fused with
// the W^{1/2} noise scaling of PDESampler::Eval (/root/reference/src/PDESampler.cpp:352-358).
//
// yarn5 (TRNG 4.19, not vendored in the reference tree; restated from the published algorithm): multiple
// recursive generator of order 5 over the prime field m = 2^31 - 1, r_n = sum_i a_i r_{n-i} mod m, default
// parameter set a = (107374182, 0, 0, 0, 104480), default state (0,1,1,1,1); output 0 if r_n == 0 else
// g^{r_n} mod m with g = 123567893.  The stream is addressed by absolute position: a thread jumps to its
// first position with precomputed powers C^(2^b) of the 5x5 companion matrix and then steps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "detmath.h"

namespace pmc {

constexpr uint32_t YARN_M = 2147483647u;
constexpr uint64_t YARN_GEN = 123567893ull;

struct RngTables {
    uint32_t jump[64][25];   // jump[b] = C^(2^b) mod m, row-major 5x5
    uint32_t powtab[4][256]; // powtab[k][j] = g^(j * 256^k) mod m
    uint32_t a[5];           // recurrence coefficients (after an optional leapfrog split)
    uint32_t r0[5];          // state at stream position 0
};

__host__ __device__ __forceinline__ uint32_t modm(uint64_t x)
{
    // x < 2^63: fold twice, then one conditional subtraction
    x = (x & YARN_M) + (x >> 31);
    x = (x & YARN_M) + (x >> 31);
    return (uint32_t)(x >= YARN_M ? x - YARN_M : x);
}
__host__ __device__ __forceinline__ uint32_t mulmod(uint32_t a, uint32_t b) { return modm((uint64_t)a * b); }

__device__ __forceinline__ void yarn5_step(uint32_t r[5], const uint32_t a[5])
{
    // each product < 2^62, the sum of five < 2^63 * 5/2: reduce the partial sums pairwise to stay below 2^63
    uint64_t t = (uint64_t)a[0] * r[0] + (uint64_t)a[1] * r[1];
    t = (uint64_t)modm(t) + (uint64_t)a[2] * r[2];
    t = (uint64_t)modm(t) + (uint64_t)a[3] * r[3];
    t = (uint64_t)modm(t) + (uint64_t)a[4] * r[4];
    r[4] = r[3]; r[3] = r[2]; r[2] = r[1]; r[1] = r[0];
    r[0] = modm(t);
}

__device__ __forceinline__ uint32_t yarn5_output(uint32_t r, const uint32_t (*pw)[256])
{
    if (r == 0) return 0;
    uint32_t v = pw[0][r & 255u];
    v = mulmod(v, pw[1][(r >> 8) & 255u]);
    v = mulmod(v, pw[2][(r >> 16) & 255u]);
    v = mulmod(v, pw[3][(r >> 24) & 255u]);
    return v;
}

__device__ __forceinline__ void yarn5_jump(uint32_t r[5], uint64_t n, const uint32_t (*J)[25])
{
    for (int b = 0; n != 0; ++b, n >>= 1) {
        if (n & 1ull) {
            uint32_t v[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                uint64_t t = 0;
#pragma unroll
                for (int k = 0; k < 5; ++k) t = (uint64_t)modm(t) + (uint64_t)J[b][i * 5 + k] * r[k];
                v[i] = modm(t);
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) r[i] = v[i];
        }
    }
}

// trng::math::inv_Phi: Acklam's rational approximation + one Halley step (restated; TRNG source absent).
// Every operation is an explicit round-to-nearest multiply/add/divide (no FMA contraction) and erf/erfc/exp/log are the
// deterministic ones of detmath.h, so a host build of the same sequence (the CPU oracle) produces bit-identical
// deviates (tests/test_gpu_parity.py::test_normal_deviates_bit_exact).
__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }

__device__ __forceinline__ double dev_Phi(double x)
{
    const double one_over_sqrt_2 = 0.70710678118654752440;
    x = dm(x, one_over_sqrt_2);
    if (x < dm(-0.6744897501960817, one_over_sqrt_2)) return dm(0.5, pmc_erfc(-x));
    if (x > dm(+0.6744897501960817, one_over_sqrt_2)) return da(1.0, -dm(0.5, pmc_erfc(x)));
    return da(0.5, dm(0.5, pmc_erf(x)));
}

__device__ __forceinline__ double dev_inv_Phi(double x)
{
    const double a0 = -3.969683028665376e+01, a1 = 2.209460984245205e+02, a2 = -2.759285104469687e+02,
                 a3 = 1.383577518672690e+02, a4 = -3.066479806614716e+01, a5 = 2.506628277459239e+00;
    const double b0 = -5.447609879822406e+01, b1 = 1.615858368580409e+02, b2 = -1.556989798598866e+02,
                 b3 = 6.680131188771972e+01, b4 = -1.328068155288572e+01;
    const double c0 = -7.784894002430293e-03, c1 = -3.223964580411365e-01, c2 = -2.400758277161838e+00,
                 c3 = -2.549732539343734e+00, c4 = 4.374664141464968e+00, c5 = 2.938163982698783e+00;
    const double d0 = 7.784695709041462e-03, d1 = 3.224671290700398e-01, d2 = 2.445134137142996e+00,
                 d3 = 3.754408661907416e+00;
    const double x_low = 0.02425, x_high = 1.0 - 0.02425;
    double t, q;
    if (x < x_low) {
        q = __dsqrt_rn(dm(-2.0, pmc_log(x)));
        const double num = da(dm(da(dm(da(dm(da(dm(da(dm(c0, q), c1), q), c2), q), c3), q), c4), q), c5);
        const double den = da(dm(da(dm(da(dm(da(dm(d0, q), d1), q), d2), q), d3), q), 1.0);
        t = __ddiv_rn(num, den);
    } else if (x < x_high) {
        q = da(x, -0.5);
        const double r = dm(q, q);
        const double num = dm(da(dm(da(dm(da(dm(da(dm(da(dm(a0, r), a1), r), a2), r), a3), r), a4), r), a5), q);
        const double den = da(dm(da(dm(da(dm(da(dm(da(dm(b0, r), b1), r), b2), r), b3), r), b4), r), 1.0);
        t = __ddiv_rn(num, den);
    } else {
        q = __dsqrt_rn(dm(-2.0, pmc_log(da(1.0, -x))));
        const double num = da(dm(da(dm(da(dm(da(dm(da(dm(c0, q), c1), q), c2), q), c3), q), c4), q), c5);
        const double den = da(dm(da(dm(da(dm(da(dm(d0, q), d1), q), d2), q), d3), q), 1.0);
        t = -__ddiv_rn(num, den);
    }
    // one step of Halley's rational method
    const double sqrt_2pi = 2.50662827463100050242;
    const double e = da(dev_Phi(t), -x);
    const double u = dm(dm(e, sqrt_2pi), pmc_exp(__ddiv_rn(dm(t, t), 2.0)));
    t = da(t, -__ddiv_rn(u, da(1.0, __ddiv_rn(dm(t, u), 2.0))));
    return t;
}

// trng::normal_dist<double>::operator()(R&): icdf(uniformoo(r)) = inv_Phi(u) * sigma + mu with u = (x + 1) / 2^31 for the
// engine output x in [0, 2^31 - 2] (utility::uniformoo<double>: the open interval (0,1)).
__device__ __forceinline__ double dev_normal_from_engine(uint32_t v, double mu, double sigma)
{
    const double u = dm((double)v + 1.0, 1.0 / 2147483648.0);
    return da(dm(dev_inv_Phi(u), sigma), mu);
}

// out[i] = normal deviate of the engine output in[i] (the map alone, for caller-chosen engine values)
__global__ void __launch_bounds__(256) k_rng_map(int64_t n, const int32_t *__restrict__ in, double *__restrict__ out, double mu,
                                                 double sigma)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = dev_normal_from_engine((uint32_t)in[i], mu, sigma);
}

// Work decomposition: thread (j, c) produces values i in [c*T, min(ni, (c+1)*T)) of sequence j, i.e. stream
// positions pos0 + j*pstride + i, and writes f(value) to out[i*si + j*sj].  MODE 0: raw int32 engine output
// (out_i); MODE 1: normal deviate mu + sigma*inv_Phi(u); MODE 2: the SPDE right-hand side
// (-g * normal) * w_sqrt[i] (PDESampler.cpp:352-358).  Values whose flat index j*pstride + i >= limit are skipped.
struct RngArgs {
    uint64_t pos0, pstride, limit;
    int64_t nj, ni, si, sj;
    int T;
    double mu, sigma, neg_g;
    const double *w_sqrt;
    double *out;
    int32_t *out_i;
};

template <int MODE>
__global__ void __launch_bounds__(128) k_rng(const RngArgs a, const RngTables *__restrict__ tab)
{
    __shared__ uint32_t J[64][25];
    __shared__ uint32_t pw[4][256];
    for (int i = threadIdx.x; i < 64 * 25; i += blockDim.x) (&J[0][0])[i] = (&tab->jump[0][0])[i];
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&pw[0][0])[i] = (&tab->powtab[0][0])[i];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t c = blockIdx.y;
    if (j >= a.nj) return;
    const int64_t i0 = c * a.T;
    if (i0 >= a.ni) return;
    const int64_t i1 = min(a.ni, i0 + (int64_t)a.T);
    uint32_t r[5], co[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { r[k] = tab->r0[k]; co[k] = tab->a[k]; }
    yarn5_jump(r, a.pos0 + (uint64_t)j * a.pstride + (uint64_t)i0, J);
    for (int64_t i = i0; i < i1; ++i) {
        yarn5_step(r, co);
        if ((uint64_t)j * a.pstride + (uint64_t)i >= a.limit) break;
        const uint32_t v = yarn5_output(r[0], pw);
        const size_t o = (size_t)(i * a.si + j * a.sj);
        if (MODE == 0) {
            a.out_i[o] = (int32_t)v;
        } else {
            double z = dev_normal_from_engine(v, a.mu, a.sigma);
            if (MODE == 2) z = dm(dm(a.neg_g, z), __ldg(a.w_sqrt + i));
            a.out[o] = z;
        }
    }
}

// ---- host side: tables ---------------------------------------------------------------------------------
inline uint32_t h_powmod(uint64_t b, uint64_t e)
{
    uint64_t r = 1;
    b %= YARN_M;
    while (e) {
        if (e & 1) r = (r * b) % YARN_M;
        b = (b * b) % YARN_M;
        e >>= 1;
    }
    return (uint32_t)r;
}

inline void h_mat5_mul(uint32_t C[25], const uint32_t A[25], const uint32_t B[25])
{
    uint32_t T[25];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            uint64_t s = 0;
            for (int k = 0; k < 5; ++k) s = (s + (uint64_t)A[i * 5 + k] * B[k * 5 + j]) % YARN_M;
            T[i * 5 + j] = (uint32_t)s;
        }
    for (int i = 0; i < 25; ++i) C[i] = T[i];
}

inline void h_yarn_jump(uint32_t r[5], const uint32_t a[5], uint64_t n)
{
    uint32_t Cm[25] = {0}, R[25] = {0};
    for (int j = 0; j < 5; ++j) Cm[j] = a[j];
    for (int i = 1; i < 5; ++i) Cm[i * 5 + i - 1] = 1;
    for (int i = 0; i < 5; ++i) R[i * 5 + i] = 1;
    while (n) {
        if (n & 1) h_mat5_mul(R, Cm, R);
        h_mat5_mul(Cm, Cm, Cm);
        n >>= 1;
    }
    uint32_t v[5];
    for (int i = 0; i < 5; ++i) {
        uint64_t s = 0;
        for (int k = 0; k < 5; ++k) s = (s + (uint64_t)R[i * 5 + k] * r[k]) % YARN_M;
        v[i] = (uint32_t)s;
    }
    for (int i = 0; i < 5; ++i) r[i] = v[i];
}

// trng::yarn5::split(s, n): leapfrog sub-stream n of s (elements n, n+s, n+2s, ... of the parent stream).
// Returns false if the 5x5 modular system is singular (cannot happen for a maximal-period generator).
inline bool h_yarn_split(uint32_t a[5], uint32_t r[5], unsigned s, unsigned n)
{
    if (s <= 1 || n >= s) return true;
    uint64_t q[10];
    h_yarn_jump(r, a, (uint64_t)n + 1);
    q[0] = r[0];
    for (int i = 1; i < 10; ++i) {
        h_yarn_jump(r, a, s);
        q[i] = r[0];
    }
    uint64_t A[5][6];
    for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < 5; ++j) A[i][j] = q[5 + i - 1 - j];
        A[i][5] = q[5 + i];
    }
    for (int c = 0; c < 5; ++c) {
        int piv = -1;
        for (int rr = c; rr < 5; ++rr)
            if (A[rr][c] != 0) { piv = rr; break; }
        if (piv < 0) return false;
        if (piv != c)
            for (int j = 0; j < 6; ++j) { uint64_t t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        const uint64_t inv = h_powmod(A[c][c], YARN_M - 2);
        for (int j = 0; j < 6; ++j) A[c][j] = (A[c][j] * inv) % YARN_M;
        for (int rr = 0; rr < 5; ++rr) {
            if (rr == c || A[rr][c] == 0) continue;
            const uint64_t f = A[rr][c];
            for (int j = 0; j < 6; ++j) A[rr][j] = (A[rr][j] + YARN_M - (f * A[c][j]) % YARN_M) % YARN_M;
        }
    }
    uint32_t na[5];
    for (int j = 0; j < 5; ++j) na[j] = (uint32_t)A[j][5];
    if (na[4] == 0) return false;
    // state = (q4, q3, q2, q1, q0), then five steps backwards with the new coefficients
    uint32_t st[5] = {(uint32_t)q[4], (uint32_t)q[3], (uint32_t)q[2], (uint32_t)q[1], (uint32_t)q[0]};
    const uint64_t inv4 = h_powmod(na[4], YARN_M - 2);
    for (int k = 0; k < 5; ++k) {
        uint64_t t = st[0];
        for (int i = 0; i < 4; ++i) t = (t + YARN_M - ((uint64_t)na[i] * st[i + 1]) % YARN_M) % YARN_M;
        t = (t * inv4) % YARN_M;
        st[0] = st[1]; st[1] = st[2]; st[2] = st[3]; st[3] = st[4]; st[4] = (uint32_t)t;
    }
    for (int i = 0; i < 5; ++i) { a[i] = na[i]; r[i] = st[i]; }
    return true;
}

inline void h_build_rng_tables(RngTables &t, int nparts, int mypart)
{
    uint32_t a[5] = {107374182u, 0u, 0u, 0u, 104480u};
    uint32_t r[5] = {0u, 1u, 1u, 1u, 1u};
    if (nparts > 1) h_yarn_split(a, r, (unsigned)nparts, (unsigned)mypart);
    for (int i = 0; i < 5; ++i) { t.a[i] = a[i]; t.r0[i] = r[i]; }
    uint32_t Cm[25] = {0};
    for (int j = 0; j < 5; ++j) Cm[j] = a[j];
    for (int i = 1; i < 5; ++i) Cm[i * 5 + i - 1] = 1;
    for (int b = 0; b < 64; ++b) {
        for (int i = 0; i < 25; ++i) t.jump[b][i] = Cm[i];
        h_mat5_mul(Cm, Cm, Cm);
    }
    for (int k = 0; k < 4; ++k) {
        const uint32_t base = h_powmod(YARN_GEN, 1ull << (8 * k));
        uint64_t v = 1;
        for (int j = 0; j < 256; ++j) {
            t.powtab[k][j] = (uint32_t)v;
            v = (v * base) % YARN_M;
        }
    }
}

}  // namespace pmc
