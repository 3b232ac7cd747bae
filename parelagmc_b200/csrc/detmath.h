/*
 * detmath.h -- deterministic FP64 exp / log / erf / erfc for the normal-deviate map of NormalDistributionSampler
 * (trng::math::inv_Phi and Phi, reached from /root/reference/src/NormalDistributionSampler.cpp:31-37).
 *
 * TRNG evaluates Phi with libm's erf/erfc and the Halley step with libm's exp; CUDA's and glibc's implementations round
 * differently in the last place, and the Halley step divides that difference by the density, so two builds that each
 * call "their" libm disagree in the last bits of ~13 % of the deviates.  This header is ONE operation sequence of
 * IEEE-754 round-to-nearest additions, multiplications and divisions (no fused multiply-add, no library call) that
 * nvcc compiles for the device (explicit __dadd_rn / __dmul_rn / __ddiv_rn, which the compiler never contracts) and a C
 * compiler compiles for the host (plain operators; build with -ffp-contract=off): both produce bit-identical doubles.
 * It is included by csrc/rng.cuh (product) and, as the same arithmetic, by the CPU checker oracle/pmc_oracle.c.
 *
 * Accuracy (tools/gen_detmath.py, tests/test_oracle.py::test_detmath_accuracy, against mpmath): exp, log <= 1 ulp,
 * erf <= 1 ulp, erfc <= 3 ulp on the domain the sampler reaches (|x| <= 6.6) -- the same class as glibc's, so the
 * deviates agree with a glibc-based evaluation to a few ulp (tests/test_oracle.py::test_normals_vs_libm).
 */
#ifndef PMC_DETMATH_H
#define PMC_DETMATH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define PMC_DM_FN __device__ __forceinline__
#define PMC_DM_TABLE static __device__ const double
#define PMC_DM_ADD(a, b) __dadd_rn((a), (b))
#define PMC_DM_SUB(a, b) __dadd_rn((a), -(b))
#define PMC_DM_MUL(a, b) __dmul_rn((a), (b))
#define PMC_DM_DIV(a, b) __ddiv_rn((a), (b))
PMC_DM_FN int64_t pmc_dm_bits(double x) { return __double_as_longlong(x); }
PMC_DM_FN double pmc_dm_from_bits(int64_t i) { return __longlong_as_double(i); }
#else
#include <string.h>
#define PMC_DM_FN static inline
#define PMC_DM_TABLE static const double
#define PMC_DM_ADD(a, b) ((a) + (b))
#define PMC_DM_SUB(a, b) ((a) - (b))
#define PMC_DM_MUL(a, b) ((a) * (b))
#define PMC_DM_DIV(a, b) ((a) / (b))
PMC_DM_FN int64_t pmc_dm_bits(double x) { int64_t i; memcpy(&i, &x, 8); return i; }
PMC_DM_FN double pmc_dm_from_bits(int64_t i) { double x; memcpy(&x, &i, 8); return x; }
#endif

#include "detmath_tables.h"

/* 2^k as a double, k in [-1022, 1023] */
PMC_DM_FN double pmc_dm_pow2(int k) { return pmc_dm_from_bits((int64_t)(k + 1023) << 52); }

/* exp(x): k = nint(x / ln 2), r = (x - k ln2_hi) - k ln2_lo, Taylor polynomial of degree 14 on |r| <= 0.35, scaling
 * by 2^k through the exponent field. */
PMC_DM_FN double pmc_exp(double x)
{
    if (x != x) return x;
    if (x > 709.78) return pmc_dm_from_bits((int64_t)0x7ff0000000000000LL);
    if (x < -745.2) return 0.0;
    const double t = PMC_DM_MUL(x, PMC_DM_INVLN2);
    const int k = (int)(t >= 0.0 ? PMC_DM_ADD(t, 0.5) : PMC_DM_SUB(t, 0.5));
    const double kd = (double)k;
    const double hi = PMC_DM_SUB(x, PMC_DM_MUL(kd, PMC_DM_LN2HI)); /* exact */
    const double r = PMC_DM_SUB(hi, PMC_DM_MUL(kd, PMC_DM_LN2LO));
    double p = pmc_dm_exp_c[PMC_DM_EXP_TERMS - 1];
    for (int i = PMC_DM_EXP_TERMS - 2; i >= 0; --i) p = PMC_DM_ADD(PMC_DM_MUL(p, r), pmc_dm_exp_c[i]);
    /* exp(r) = 1 + (r + r^2 p) */
    const double e = PMC_DM_ADD(1.0, PMC_DM_ADD(r, PMC_DM_MUL(PMC_DM_MUL(r, r), p)));
    if (k > 1023) return PMC_DM_MUL(PMC_DM_MUL(e, pmc_dm_pow2(1023)), pmc_dm_pow2(k - 1023));
    if (k < -1022) return PMC_DM_MUL(PMC_DM_MUL(e, pmc_dm_pow2(k + 1000)), pmc_dm_pow2(-1000));
    return PMC_DM_MUL(e, pmc_dm_pow2(k));
}

/* log(x), x > 0: x = m 2^e with m in [sqrt(1/2), sqrt 2), s = (m - 1) / (m + 1), log m = 2 s + 2 s z (1/3 + z/5 + ...) */
PMC_DM_FN double pmc_log(double x)
{
    if (x != x) return x;
    if (x < 0.0) return pmc_dm_from_bits((int64_t)0x7ff8000000000000LL);
    if (x == 0.0) return pmc_dm_from_bits((int64_t)0xfff0000000000000LL);
    int e = 0;
    int64_t b = pmc_dm_bits(x);
    if (b >= (int64_t)0x7ff0000000000000LL) return x;
    if (b < (int64_t)0x0010000000000000LL) { /* subnormal */
        x = PMC_DM_MUL(x, 18014398509481984.0);  /* 2^54 */
        e = -54;
        b = pmc_dm_bits(x);
    }
    e += (int)(b >> 52) - 1023;
    b = (b & (int64_t)0x000fffffffffffffLL) | (int64_t)0x3ff0000000000000LL;
    double m = pmc_dm_from_bits(b);
    if (m > 1.4142135623730951) {
        m = PMC_DM_MUL(m, 0.5);
        e += 1;
    }
    const double f = PMC_DM_SUB(m, 1.0); /* exact */
    const double s = PMC_DM_DIV(f, PMC_DM_ADD(2.0, f));
    const double z = PMC_DM_MUL(s, s);
    double p = pmc_dm_log_c[PMC_DM_LOG_TERMS - 1];
    for (int i = PMC_DM_LOG_TERMS - 2; i >= 0; --i) p = PMC_DM_ADD(PMC_DM_MUL(p, z), pmc_dm_log_c[i]);
    const double lm = PMC_DM_ADD(PMC_DM_MUL(2.0, s), PMC_DM_MUL(PMC_DM_MUL(s, z), p));
    const double ed = (double)e;
    return PMC_DM_ADD(PMC_DM_MUL(ed, PMC_DM_LN2HI), PMC_DM_ADD(PMC_DM_MUL(ed, PMC_DM_LN2LO), lm));
}

/* erf on |x| < 0.5: x + x (2/sqrt(pi) - 1 + z (c1 + z (c2 + ...))), z = x^2 */
PMC_DM_FN double pmc_dm_erf_small(double x)
{
    const double z = PMC_DM_MUL(x, x);
    double p = pmc_dm_erf_c[PMC_DM_ERF_TERMS - 1];
    for (int i = PMC_DM_ERF_TERMS - 2; i >= 0; --i) p = PMC_DM_ADD(PMC_DM_MUL(p, z), pmc_dm_erf_c[i]);
    const double y = PMC_DM_ADD(PMC_DM_ERF_EFX, PMC_DM_MUL(z, p));
    return PMC_DM_ADD(x, PMC_DM_MUL(x, y));
}

/* erfc on x >= 0.375: erfcx(x) exp(-x^2).  erfcx from a table of Taylor expansions (intervals of width 1/4, degree 17)
 * below 6.875 and from its asymptotic series above; exp(-x^2) with x split into a 26-bit head z (z^2 exact) and a tail:
 * exp(-z^2) (1 + d + d^2/2 + d^3/6), d = (z - x)(z + x), |d| < 3e-5. */
PMC_DM_FN double pmc_dm_erfc_large(double x)
{
    if (x > 27.3) return 0.0;
    double cx;
    const double u = PMC_DM_MUL(PMC_DM_SUB(x, PMC_DM_ERFCX_X0), 4.0);
    const int i = (int)u;
    if (i < PMC_DM_ERFCX_N) {
        const double ctr = PMC_DM_ADD(PMC_DM_ERFCX_X0 + 0.125, PMC_DM_MUL((double)i, 0.25)); /* exact */
        const double h = PMC_DM_SUB(x, ctr);                                               /* exact */
        const double *c = pmc_dm_erfcx_t + i * (PMC_DM_ERFCX_DEG + 1);
        cx = c[PMC_DM_ERFCX_DEG];
        for (int j = PMC_DM_ERFCX_DEG - 1; j >= 0; --j) cx = PMC_DM_ADD(PMC_DM_MUL(cx, h), c[j]);
    } else {
        const double w = PMC_DM_DIV(1.0, PMC_DM_MUL(x, x));
        double p = pmc_dm_asym_c[PMC_DM_ASYM_TERMS - 1];
        for (int j = PMC_DM_ASYM_TERMS - 2; j >= 0; --j) p = PMC_DM_ADD(PMC_DM_MUL(p, w), pmc_dm_asym_c[j]);
        cx = PMC_DM_DIV(PMC_DM_MUL(PMC_DM_INV_SQRT_PI, PMC_DM_ADD(1.0, PMC_DM_MUL(w, p))), x);
    }
    const double z = pmc_dm_from_bits(pmc_dm_bits(x) & (int64_t)0xfffffffff8000000LL);
    const double d = PMC_DM_MUL(PMC_DM_SUB(z, x), PMC_DM_ADD(z, x));
    const double ez = pmc_exp(-PMC_DM_MUL(z, z));
    const double corr = PMC_DM_ADD(1.0, PMC_DM_MUL(d, PMC_DM_ADD(1.0, PMC_DM_MUL(d, PMC_DM_ADD(0.5, PMC_DM_MUL(d, 0x1.5555555555555p-3))))));
    return PMC_DM_MUL(PMC_DM_MUL(ez, corr), cx);
}

PMC_DM_FN double pmc_erf(double x)
{
    if (x != x) return x;
    const double ax = x < 0.0 ? -x : x;
    if (ax < 0.5) return pmc_dm_erf_small(x);
    const double r = PMC_DM_SUB(1.0, pmc_dm_erfc_large(ax));
    return x < 0.0 ? -r : r;
}

PMC_DM_FN double pmc_erfc(double x)
{
    if (x != x) return x;
    if (x < 0.0) {
        if (x > -0.5) return PMC_DM_SUB(1.0, pmc_dm_erf_small(x));
        return PMC_DM_SUB(2.0, pmc_dm_erfc_large(-x));
    }
    if (x < 0.5) return PMC_DM_SUB(1.0, pmc_dm_erf_small(x));
    return pmc_dm_erfc_large(x);
}

#endif /* PMC_DETMATH_H */
