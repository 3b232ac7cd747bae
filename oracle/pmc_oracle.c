/*
 * pmc_oracle.c -- CPU ORACLE (test infrastructure only; see pmc_oracle.h for scope and parity status).
 *
 * Restates, in plain C, the per-sample hot path of ParELAGMC:
 *   NormalDistributionSampler   /root/reference/src/NormalDistributionSampler.cpp:17-37
 *   PDESampler::Sample/Eval     /root/reference/src/PDESampler.cpp:336-535
 *   DarcySolver::SolveFwd       /root/reference/src/DarcySolver.cpp:416-437,472-520,562-649
 *   MLMC_Manager::InitRun       /root/reference/src/MLMC_Manager.cpp:103-179
 *   computeNSamplesMSE          /root/reference/src/MLMC_Manager.cpp:300-401, src/MC_Manager.cpp:194-239
 *   expWRegression              /root/reference/src/Utilities.cpp:257-283
 * The linear solver is MFEM-style preconditioned MINRES with the reference's "MINRES-BJ-GS" structure
 * (block-Jacobi: symmetric Gauss-Seidel sweeps on M, one multigrid V-cycle on B diag(M)^-1 B^T,
 * /root/reference/examples/example_helpers/CreateMLMCParameterList.hpp:58-118); BoomerAMG itself is
 * replaced by a V-cycle on the hierarchy's own piecewise-constant prolongators with Galerkin coarse
 * matrices rebuilt per sample (as the reference rebuilds its AMG per sample, DarcySolver.cpp:596-601).
 * At tolerance 1e-12 the solution does not depend on the preconditioner.
 */
#include "pmc_oracle.h"
/* One operation sequence for exp/log/erf/erfc shared with the CUDA build, so that the deviates are bit-identical on
 * both sides (libm's and CUDA's own implementations differ in the last place); see the header for accuracy. */
#include "../parelagmc_b200/csrc/detmath.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ============================================================================================ */
/* yarn5                                                                                        */
/* ============================================================================================ */
#define PO_M 2147483647ULL /* 2^31 - 1 */
#define PO_GEN 123567893ULL

static inline uint64_t mulmod(uint64_t a, uint64_t b) { return (a * b) % PO_M; }

static uint64_t powmod(uint64_t b, uint64_t e)
{
    uint64_t r = 1;
    b %= PO_M;
    while (e) {
        if (e & 1) r = mulmod(r, b);
        b = mulmod(b, b);
        e >>= 1;
    }
    return r;
}

void po_yarn5_init(po_yarn5 *g)
{
    /* parameter set "LEcuyer1", default status (0,1,1,1,1) */
    g->a[0] = 107374182; g->a[1] = 0; g->a[2] = 0; g->a[3] = 0; g->a[4] = 104480;
    g->r[0] = 0; g->r[1] = 1; g->r[2] = 1; g->r[3] = 1; g->r[4] = 1;
}

static inline void yarn5_step(po_yarn5 *g)
{
    uint64_t t = 0;
    for (int i = 0; i < 5; ++i) t = (t + mulmod((uint64_t)g->a[i], (uint64_t)g->r[i])) % PO_M;
    g->r[4] = g->r[3]; g->r[3] = g->r[2]; g->r[2] = g->r[1]; g->r[1] = g->r[0];
    g->r[0] = (int32_t)t;
}

int32_t po_yarn5_next(po_yarn5 *g)
{
    yarn5_step(g);
    if (g->r[0] == 0) return 0;
    return (int32_t)powmod(PO_GEN, (uint64_t)g->r[0]);
}

static void mat5_mul(uint64_t C[25], const uint64_t A[25], const uint64_t B[25])
{
    uint64_t T[25];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            uint64_t s = 0;
            for (int k = 0; k < 5; ++k) s = (s + mulmod(A[i * 5 + k], B[k * 5 + j])) % PO_M;
            T[i * 5 + j] = s;
        }
    memcpy(C, T, sizeof T);
}

void po_yarn5_jump(po_yarn5 *g, uint64_t n)
{
    if (n < 16) {
        for (uint64_t i = 0; i < n; ++i) yarn5_step(g);
        return;
    }
    uint64_t C[25] = {0}, R[25] = {0};
    for (int j = 0; j < 5; ++j) C[j] = (uint64_t)g->a[j];
    for (int i = 1; i < 5; ++i) C[i * 5 + (i - 1)] = 1;
    for (int i = 0; i < 5; ++i) R[i * 5 + i] = 1;
    while (n) {
        if (n & 1) mat5_mul(R, C, R);
        mat5_mul(C, C, C);
        n >>= 1;
    }
    uint64_t v[5];
    for (int i = 0; i < 5; ++i) {
        uint64_t s = 0;
        for (int k = 0; k < 5; ++k) s = (s + mulmod(R[i * 5 + k], (uint64_t)g->r[k])) % PO_M;
        v[i] = s;
    }
    for (int i = 0; i < 5; ++i) g->r[i] = (int32_t)v[i];
}

static uint64_t invmod(uint64_t a) { return powmod(a, PO_M - 2); }

static int yarn5_backward(po_yarn5 *g)
{
    /* r0 = a0 r1 + a1 r2 + a2 r3 + a3 r4 + a4 x  ->  x */
    if (g->a[4] == 0) return -1;
    uint64_t t = (uint64_t)g->r[0];
    for (int i = 0; i < 4; ++i)
        t = (t + PO_M - mulmod((uint64_t)g->a[i], (uint64_t)g->r[i + 1])) % PO_M;
    t = mulmod(t, invmod((uint64_t)g->a[4]));
    g->r[0] = g->r[1]; g->r[1] = g->r[2]; g->r[2] = g->r[3]; g->r[3] = g->r[4];
    g->r[4] = (int32_t)t;
    return 0;
}

void po_yarn5_split(po_yarn5 *g, unsigned s, unsigned n)
{
    /* leapfrog: sub-stream n of s takes elements n, n+s, n+2s, ... of the original stream */
    if (s <= 1 || n >= s) return;
    uint64_t q[10];
    po_yarn5_jump(g, (uint64_t)n + 1);
    q[0] = (uint64_t)g->r[0];
    for (int i = 1; i < 10; ++i) {
        po_yarn5_jump(g, s);
        q[i] = (uint64_t)g->r[0];
    }
    /* solve q[i] = sum_j b[j] q[i-1-j], i = 5..9, over GF(m) */
    uint64_t A[5][6];
    for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < 5; ++j) A[i][j] = q[5 + i - 1 - j];
        A[i][5] = q[5 + i];
    }
    for (int c = 0; c < 5; ++c) {
        int piv = -1;
        for (int r = c; r < 5; ++r)
            if (A[r][c] != 0) { piv = r; break; }
        if (piv < 0) return; /* singular: cannot happen for a maximal-period MRG */
        if (piv != c)
            for (int j = 0; j < 6; ++j) { uint64_t t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        uint64_t inv = invmod(A[c][c]);
        for (int j = 0; j < 6; ++j) A[c][j] = mulmod(A[c][j], inv);
        for (int r = 0; r < 5; ++r) {
            if (r == c || A[r][c] == 0) continue;
            uint64_t f = A[r][c];
            for (int j = 0; j < 6; ++j) A[r][j] = (A[r][j] + PO_M - mulmod(f, A[c][j])) % PO_M;
        }
    }
    for (int j = 0; j < 5; ++j) g->a[j] = (int32_t)A[j][5];
    g->r[0] = (int32_t)q[4]; g->r[1] = (int32_t)q[3]; g->r[2] = (int32_t)q[2];
    g->r[3] = (int32_t)q[1]; g->r[4] = (int32_t)q[0];
    for (int i = 0; i < 5; ++i) yarn5_backward(g);
}

void po_yarn5_fill_int(po_yarn5 *g, int64_t n, int32_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = po_yarn5_next(g);
}

/* utility::uniformoo<double>: (x - min + 1) / (max - min + 2) with min = 0, max = 2^31 - 2 */
double po_uniformoo(int32_t x) { return ((double)x + 1.0) * (1.0 / 2147483648.0); }

/* trng::math::Phi / inv_Phi (TRNG 4.19 special_functions.hpp; restated, source absent): Acklam's rational approximation
 * followed by one step of Halley's rational method.  `libm` selects glibc's erf/erfc/exp/log (what a TRNG build on this
 * host would call) instead of the deterministic ones of detmath.h; the product and the parity tests use libm = 0. */
static double po_Phi_impl(double x, int libm)
{
    const double one_over_sqrt_2 = 0.70710678118654752440;
    x *= one_over_sqrt_2;
    if (x < -0.6744897501960817 * one_over_sqrt_2) return 0.5 * (libm ? erfc(-x) : pmc_erfc(-x));
    if (x > +0.6744897501960817 * one_over_sqrt_2) return 1.0 - 0.5 * (libm ? erfc(x) : pmc_erfc(x));
    return 0.5 + 0.5 * (libm ? erf(x) : pmc_erf(x));
}

static double po_inv_Phi_approx(double x, int libm)
{
    /* P. J. Acklam's rational approximation */
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00};
    const double x_low = 0.02425, x_high = 1.0 - 0.02425;
    if (x < 0.0 || x > 1.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == 1.0) return INFINITY;
    double t, q;
    if (x < x_low) {
        q = sqrt(-2.0 * (libm ? log(x) : pmc_log(x)));
        t = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    } else if (x < x_high) {
        q = x - 0.5;
        double r = q * q;
        t = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
    } else {
        q = sqrt(-2.0 * (libm ? log(1.0 - x) : pmc_log(1.0 - x)));
        t = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    return t;
}

static double po_inv_Phi_impl(double x, int libm)
{
    double y = po_inv_Phi_approx(x, libm);
    if (isfinite(y)) { /* one step of Halley's rational method */
        const double sqrt_2pi = 2.50662827463100050242;
        double e = po_Phi_impl(y, libm) - x;
        double u = e * sqrt_2pi * (libm ? exp(y * y / 2.0) : pmc_exp(y * y / 2.0));
        y -= u / (1.0 + y * u / 2.0);
    }
    return y;
}

double po_inv_Phi(double x) { return po_inv_Phi_impl(x, 0); }
double po_inv_Phi_libm(double x) { return po_inv_Phi_impl(x, 1); }

/* the deterministic elementary functions on their own (accuracy tests against mpmath / libm) */
double po_det_exp(double x) { return pmc_exp(x); }
double po_det_log(double x) { return pmc_log(x); }
double po_det_erf(double x) { return pmc_erf(x); }
double po_det_erfc(double x) { return pmc_erfc(x); }
void po_det_eval(int which, int64_t n, const double *x, double *y)
{
    for (int64_t i = 0; i < n; ++i)
        y[i] = which == 0 ? pmc_exp(x[i]) : which == 1 ? pmc_log(x[i]) : which == 2 ? pmc_erf(x[i])
             : which == 3 ? pmc_erfc(x[i]) : which == 4 ? po_inv_Phi_impl(x[i], 0) : po_inv_Phi_impl(x[i], 1);
}

void po_normal_fill(po_yarn5 *g, double mu, double sigma, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = po_inv_Phi(po_uniformoo(po_yarn5_next(g))) * sigma + mu;
}

/* the map of trng::normal_dist<double> on caller-chosen engine outputs (reaches both extreme tails) */
void po_normal_map(int64_t n, const int32_t *engine, double mu, double sigma, double *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = po_inv_Phi(po_uniformoo(engine[i])) * sigma + mu;
}

/* ============================================================================================ */
/* small CSR toolkit                                                                            */
/* ============================================================================================ */
typedef struct {
    int rows, cols;
    int *rowptr, *col;
    double *val;
} csr_t;

static void csr_free(csr_t *A)
{
    free(A->rowptr); free(A->col); free(A->val);
    memset(A, 0, sizeof *A);
}

static csr_t csr_copy_in(int rows, int cols, const int *rowptr, const int *col, const double *val)
{
    csr_t A;
    A.rows = rows; A.cols = cols;
    int nnz = rowptr[rows];
    A.rowptr = (int *)malloc(sizeof(int) * (size_t)(rows + 1));
    A.col = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    A.val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    memcpy(A.rowptr, rowptr, sizeof(int) * (size_t)(rows + 1));
    memcpy(A.col, col, sizeof(int) * (size_t)nnz);
    if (val) memcpy(A.val, val, sizeof(double) * (size_t)nnz);
    else for (int i = 0; i < nnz; ++i) A.val[i] = 0.0;
    return A;
}

static csr_t csr_transpose(const csr_t *A)
{
    csr_t T;
    T.rows = A->cols; T.cols = A->rows;
    int nnz = A->rowptr[A->rows];
    T.rowptr = (int *)calloc((size_t)(T.rows + 1), sizeof(int));
    T.col = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    T.val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    for (int i = 0; i < nnz; ++i) T.rowptr[A->col[i] + 1]++;
    for (int i = 0; i < T.rows; ++i) T.rowptr[i + 1] += T.rowptr[i];
    int *pos = (int *)malloc(sizeof(int) * (size_t)(T.rows > 0 ? T.rows : 1));
    memcpy(pos, T.rowptr, sizeof(int) * (size_t)T.rows);
    for (int i = 0; i < A->rows; ++i)
        for (int p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) {
            int q = pos[A->col[p]]++;
            T.col[q] = i; T.val[q] = A->val[p];
        }
    free(pos);
    return T;
}

/* C = A * B (numeric + symbolic, sorted columns) */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

static csr_t csr_matmul(const csr_t *A, const csr_t *B)
{
    csr_t C;
    C.rows = A->rows; C.cols = B->cols;
    C.rowptr = (int *)calloc((size_t)(C.rows + 1), sizeof(int));
    int *mark = (int *)malloc(sizeof(int) * (size_t)(C.cols > 0 ? C.cols : 1));
    double *acc = (double *)calloc((size_t)(C.cols > 0 ? C.cols : 1), sizeof(double));
    for (int j = 0; j < C.cols; ++j) mark[j] = -1;
    int cap = A->rowptr[A->rows] + B->rowptr[B->rows] + 16, nnz = 0;
    C.col = (int *)malloc(sizeof(int) * (size_t)cap);
    C.val = (double *)malloc(sizeof(double) * (size_t)cap);
    for (int i = 0; i < A->rows; ++i) {
        int start = nnz;
        for (int p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) {
            int k = A->col[p];
            double a = A->val[p];
            for (int q = B->rowptr[k]; q < B->rowptr[k + 1]; ++q) {
                int j = B->col[q];
                if (mark[j] < start) {
                    if (nnz == cap) {
                        cap *= 2;
                        C.col = (int *)realloc(C.col, sizeof(int) * (size_t)cap);
                        C.val = (double *)realloc(C.val, sizeof(double) * (size_t)cap);
                    }
                    mark[j] = nnz; C.col[nnz++] = j; acc[j] = a * B->val[q];
                } else acc[j] += a * B->val[q];
            }
        }
        qsort(C.col + start, (size_t)(nnz - start), sizeof(int), cmp_int);
        for (int p = start; p < nnz; ++p) { C.val[p] = acc[C.col[p]]; mark[C.col[p]] = -1; }
        /* marks reset to -1 < any later start */
        C.rowptr[i + 1] = nnz;
    }
    free(mark); free(acc);
    return C;
}

static void csr_mult(const csr_t *A, const double *x, double *y) /* y = A x */
{
    for (int i = 0; i < A->rows; ++i) {
        double s = 0.0;
        for (int p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) s += A->val[p] * x[A->col[p]];
        y[i] = s;
    }
}

static void csr_mult_add(const csr_t *A, const double *x, double *y) /* y += A x */
{
    for (int i = 0; i < A->rows; ++i) {
        double s = 0.0;
        for (int p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) s += A->val[p] * x[A->col[p]];
        y[i] += s;
    }
}

static void csr_mult_transpose(const csr_t *A, const double *x, double *y) /* y = A^T x */
{
    for (int j = 0; j < A->cols; ++j) y[j] = 0.0;
    for (int i = 0; i < A->rows; ++i) {
        double xi = x[i];
        for (int p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) y[A->col[p]] += A->val[p] * xi;
    }
}

static double csr_diag_entry(const csr_t *A, int i)
{
    for (int p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p)
        if (A->col[p] == i) return A->val[p];
    return 0.0;
}

static double vdot(int n, const double *a, const double *b)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* ============================================================================================ */
/* preconditioner pieces                                                                        */
/* ============================================================================================ */
/* `sweeps` symmetric Gauss-Seidel sweeps on A z = r from z = 0 (hypre "L1 Gauss-Seidel", Sweeps = 3,
 * CreateMLMCParameterList.hpp:85-98; on one rank the l1 term vanishes) */
static void sym_gs(const csr_t *A, const double *r, double *z, int sweeps, int zero_init)
{
    int n = A->rows;
    if (zero_init) for (int i = 0; i < n; ++i) z[i] = 0.0;
    for (int s = 0; s < sweeps; ++s) {
        for (int i = 0; i < n; ++i) {
            double t = r[i], d = 1.0;
            for (int p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) {
                int j = A->col[p];
                if (j == i) d = A->val[p]; else t -= A->val[p] * z[j];
            }
            z[i] = t / d;
        }
        for (int i = n - 1; i >= 0; --i) {
            double t = r[i], d = 1.0;
            for (int p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) {
                int j = A->col[p];
                if (j == i) d = A->val[p]; else t -= A->val[p] * z[j];
            }
            z[i] = t / d;
        }
    }
}

#define PO_MAX_MG 16
#define PO_DENSE_MAX 600

typedef struct {
    int nlev;
    csr_t S[PO_MAX_MG];      /* S[0] fine ... */
    const csr_t *P[PO_MAX_MG]; /* borrowed prolongators P[m]: n_m x n_{m+1} */
    double *chol;            /* dense Cholesky factor of the coarsest S (row-major lower) or NULL */
    double *r[PO_MAX_MG], *x[PO_MAX_MG], *t[PO_MAX_MG];
} mg_t;

static void mg_free(mg_t *mg)
{
    for (int m = 0; m < mg->nlev; ++m) {
        csr_free(&mg->S[m]);
        free(mg->r[m]); free(mg->x[m]); free(mg->t[m]);
    }
    free(mg->chol);
    memset(mg, 0, sizeof *mg);
}

/* takes ownership of S0 */
static void mg_setup(mg_t *mg, csr_t S0, const csr_t *const *P, int nP)
{
    memset(mg, 0, sizeof *mg);
    mg->nlev = nP + 1;
    mg->S[0] = S0;
    for (int m = 0; m < nP; ++m) {
        mg->P[m] = P[m];
        csr_t Pt = csr_transpose(P[m]);
        csr_t SP = csr_matmul(&mg->S[m], P[m]);
        mg->S[m + 1] = csr_matmul(&Pt, &SP);
        csr_free(&Pt); csr_free(&SP);
    }
    for (int m = 0; m < mg->nlev; ++m) {
        size_t n = (size_t)mg->S[m].rows;
        mg->r[m] = (double *)malloc(sizeof(double) * n);
        mg->x[m] = (double *)malloc(sizeof(double) * n);
        mg->t[m] = (double *)malloc(sizeof(double) * n);
    }
    const csr_t *Sc = &mg->S[mg->nlev - 1];
    int n = Sc->rows;
    if (n <= PO_DENSE_MAX) {
        double *L = (double *)calloc((size_t)n * (size_t)n, sizeof(double));
        for (int i = 0; i < n; ++i)
            for (int p = Sc->rowptr[i]; p < Sc->rowptr[i + 1]; ++p)
                if (Sc->col[p] <= i) L[(size_t)i * n + Sc->col[p]] = Sc->val[p];
        int ok = 1;
        for (int j = 0; j < n && ok; ++j) {
            double d = L[(size_t)j * n + j];
            for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
            if (!(d > 0.0)) { ok = 0; break; }
            d = sqrt(d);
            L[(size_t)j * n + j] = d;
            for (int i = j + 1; i < n; ++i) {
                double s = L[(size_t)i * n + j];
                for (int k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
                L[(size_t)i * n + j] = s / d;
            }
        }
        if (ok) mg->chol = L; else free(L);
    }
}

static void mg_vcycle(mg_t *mg, int m, const double *r, double *x)
{
    const csr_t *S = &mg->S[m];
    int n = S->rows;
    if (m == mg->nlev - 1) {
        if (mg->chol) {
            const double *L = mg->chol;
            for (int i = 0; i < n; ++i) {
                double s = r[i];
                for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * x[k];
                x[i] = s / L[(size_t)i * n + i];
            }
            for (int i = n - 1; i >= 0; --i) {
                double s = x[i];
                for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * x[k];
                x[i] = s / L[(size_t)i * n + i];
            }
        } else {
            sym_gs(S, r, x, 20, 1);
        }
        return;
    }
    sym_gs(S, r, x, 1, 1);
    double *res = mg->t[m];
    csr_mult(S, x, res);
    for (int i = 0; i < n; ++i) res[i] = r[i] - res[i];
    csr_mult_transpose(mg->P[m], res, mg->r[m + 1]);
    mg_vcycle(mg, m + 1, mg->r[m + 1], mg->x[m + 1]);
    csr_mult_add(mg->P[m], mg->x[m + 1], x);
    sym_gs(S, r, x, 1, 0);
}

/* ============================================================================================ */
/* saddle-point system and MINRES                                                               */
/* ============================================================================================ */
typedef struct {
    int Nf, Ne;
    const csr_t *M;   /* Nf x Nf */
    const csr_t *B;   /* Ne x Nf */
    const csr_t *Bt;  /* Nf x Ne */
    const double *c;  /* (1,1) diagonal block (-alpha W) or NULL */
    mg_t *mg;         /* V-cycle for the Schur complement approximation */
    int gs_sweeps;
} saddle_t;

static void saddle_mult(const saddle_t *A, const double *x, double *y)
{
    csr_mult(A->M, x, y);
    csr_mult_add(A->Bt, x + A->Nf, y);
    csr_mult(A->B, x, y + A->Nf);
    if (A->c)
        for (int i = 0; i < A->Ne; ++i) y[A->Nf + i] += A->c[i] * x[A->Nf + i];
}

static void saddle_prec(const saddle_t *A, const double *r, double *z)
{
    sym_gs(A->M, r, z, A->gs_sweeps, 1);
    mg_vcycle(A->mg, 0, r + A->Nf, z + A->Nf);
}

/* mfem::MINRESSolver::Mult (preconditioned form; convergence on the preconditioned residual norm
 * |eta| <= max(rel_tol * eta0, abs_tol)); iterative != 0 uses x as the initial guess. */
static int minres(const saddle_t *A, const double *b, double *x, int iterative, double rel, double abs_,
                  int maxit)
{
    int n = A->Nf + A->Ne, it = 0;
    double *v0 = (double *)calloc((size_t)n, sizeof(double));
    double *v1 = (double *)calloc((size_t)n, sizeof(double));
    double *w0 = (double *)calloc((size_t)n, sizeof(double));
    double *w1 = (double *)calloc((size_t)n, sizeof(double));
    double *q = (double *)calloc((size_t)n, sizeof(double));
    double *u1 = (double *)calloc((size_t)n, sizeof(double));
    double alpha, beta, delta, rho1, rho2, rho3, eta, norm_goal;
    double gamma0 = 1.0, gamma1 = 1.0, sigma0 = 0.0, sigma1 = 0.0;
    if (iterative) {
        saddle_mult(A, x, v1);
        for (int i = 0; i < n; ++i) v1[i] = b[i] - v1[i];
    } else {
        for (int i = 0; i < n; ++i) { v1[i] = b[i]; x[i] = 0.0; }
    }
    saddle_prec(A, v1, u1);
    eta = beta = sqrt(vdot(n, u1, v1));
    norm_goal = fmax(rel * eta, abs_);
    if (eta <= norm_goal) goto done;
    for (it = 1; it <= maxit; ++it) {
        double ib = 1.0 / beta;
        for (int i = 0; i < n; ++i) { v1[i] *= ib; u1[i] *= ib; }
        saddle_mult(A, u1, q);
        alpha = vdot(n, u1, q);
        if (it > 1)
            for (int i = 0; i < n; ++i) q[i] -= beta * v0[i];
        for (int i = 0; i < n; ++i) v0[i] = q[i] - alpha * v1[i];
        delta = gamma1 * alpha - gamma0 * sigma1 * beta;
        rho3 = sigma0 * beta;
        rho2 = sigma1 * alpha + gamma0 * gamma1 * beta;
        saddle_prec(A, v0, q);
        beta = sqrt(fmax(vdot(n, v0, q), 0.0));
        rho1 = hypot(delta, beta);
        if (it == 1)
            for (int i = 0; i < n; ++i) w0[i] = u1[i] / rho1;
        else if (it == 2)
            for (int i = 0; i < n; ++i) w0[i] = u1[i] / rho1 - (rho2 / rho1) * w1[i];
        else
            for (int i = 0; i < n; ++i)
                w0[i] = (-rho3 / rho1) * w0[i] - (rho2 / rho1) * w1[i] + u1[i] / rho1;
        gamma0 = gamma1;
        gamma1 = delta / rho1;
        for (int i = 0; i < n; ++i) x[i] += gamma1 * eta * w0[i];
        sigma0 = sigma1;
        sigma1 = beta / rho1;
        eta = -sigma1 * eta;
        if (fabs(eta) <= norm_goal) break;
        { double *t = u1; u1 = q; q = t; }
        { double *t = v0; v0 = v1; v1 = t; }
        { double *t = w0; w0 = w1; w1 = t; }
    }
done:
    free(v0); free(v1); free(w0); free(w1); free(q); free(u1);
    return it > maxit ? maxit : it;
}

/* S = c_diag + B diag(d)^-1 B^T  (c_diag may be NULL) */
static csr_t schur_diag(const csr_t *B, const csr_t *Bt, const double *Mdiag, const double *cdiag)
{
    csr_t Bs = csr_copy_in(B->rows, B->cols, B->rowptr, B->col, B->val);
    for (int i = 0; i < B->rows; ++i)
        for (int p = B->rowptr[i]; p < B->rowptr[i + 1]; ++p) Bs.val[p] /= Mdiag[B->col[p]];
    csr_t S = csr_matmul(&Bs, Bt);
    csr_free(&Bs);
    if (cdiag)
        for (int i = 0; i < S.rows; ++i)
            for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p)
                if (S.col[p] == i) S.val[p] += cdiag[i];
    return S;
}

/* ============================================================================================ */
/* problem                                                                                      */
/* ============================================================================================ */
typedef struct {
    int set, Ne, Nf, lognormal;
    csr_t M, B, Bt, P, Pt;
    int hasP;
    double *Wdiag, *w_sqrt, *negaW; /* -alpha W */
    double alpha, g;
    mg_t mg;
    int mg_ready;
    /* optional field transfer to the forward problem's mesh: s = tscale .* (T field)
     * (EmbeddedPDESampler.cpp:426-435: T = meshP, no scale; L2ProjectionPDESampler.cpp:595-611: T = G^T, 1/diag(W)) */
    int hasT, n_out;
    csr_t T;
    double *tscale;
} sampler_level_t;

typedef struct {
    int set, Ne, Nf;
    int *elem_ptr, *elem_dofs;
    double *elem_mat;
    long *elem_mat_ptr;
    csr_t Mpat;      /* pattern of the assembled mass matrix */
    int *scatter;    /* for each element-matrix entry, its position in Mpat.val */
    csr_t B;         /* un-eliminated */
    csr_t Be, Bet;   /* essential columns removed (pattern kept, values zeroed) */
    int *ess_u;
    double *ess_data, *rhs, *obs;
    csr_t Pp;
    int hasP;
    /* BayesianInverseProblem: m pressure functionals (un-normalised), observed data, noise variance */
    int n_obs;
    double *gobs_func, *Gobs, noise;
} darcy_level_t;

struct po_problem {
    int nlevels;
    sampler_level_t *s;
    darcy_level_t *d;
    double rel, abs_;
    int maxit;
};

po_problem *po_create(int nlevels)
{
    po_problem *p = (po_problem *)calloc(1, sizeof *p);
    p->nlevels = nlevels;
    p->s = (sampler_level_t *)calloc((size_t)nlevels, sizeof *p->s);
    p->d = (darcy_level_t *)calloc((size_t)nlevels, sizeof *p->d);
    p->rel = 1e-6; p->abs_ = 1e-12; p->maxit = 300; /* CreateMLMCParameterList.hpp:67-69 */
    return p;
}

void po_destroy(po_problem *p)
{
    if (!p) return;
    for (int l = 0; l < p->nlevels; ++l) {
        sampler_level_t *s = &p->s[l];
        if (s->set) {
            csr_free(&s->M); csr_free(&s->B); csr_free(&s->Bt);
            if (s->hasP) { csr_free(&s->P); csr_free(&s->Pt); }
            free(s->Wdiag); free(s->w_sqrt); free(s->negaW);
            if (s->mg_ready) mg_free(&s->mg);
            if (s->hasT) { csr_free(&s->T); free(s->tscale); }
        }
        darcy_level_t *d = &p->d[l];
        if (d->set) {
            free(d->elem_ptr); free(d->elem_dofs); free(d->elem_mat); free(d->elem_mat_ptr);
            csr_free(&d->Mpat); free(d->scatter);
            csr_free(&d->B); csr_free(&d->Be); csr_free(&d->Bet);
            free(d->ess_u); free(d->ess_data); free(d->rhs); free(d->obs);
            if (d->hasP) csr_free(&d->Pp);
            free(d->gobs_func); free(d->Gobs);
        }
    }
    free(p->s); free(p->d); free(p);
}

void po_set_tolerances(po_problem *p, double rel, double abs_, int maxit)
{
    p->rel = rel; p->abs_ = abs_; p->maxit = maxit;
}

static double *dup_d(const double *x, size_t n)
{
    double *y = (double *)malloc(sizeof(double) * (n ? n : 1));
    if (x) memcpy(y, x, sizeof(double) * n); else memset(y, 0, sizeof(double) * n);
    return y;
}

int po_set_sampler_level(po_problem *p, int level, int Ne, int Nf,
                         const int *M_rowptr, const int *M_col, const double *M_val,
                         const int *B_rowptr, const int *B_col, const double *B_val,
                         const double *Wdiag,
                         int P_cols, const int *P_rowptr, const int *P_col, const double *P_val,
                         double alpha, double matern_coeff, int lognormal)
{
    if (level < 0 || level >= p->nlevels) return -1;
    sampler_level_t *s = &p->s[level];
    s->Ne = Ne; s->Nf = Nf; s->lognormal = lognormal; s->alpha = alpha; s->g = matern_coeff;
    s->M = csr_copy_in(Nf, Nf, M_rowptr, M_col, M_val);
    s->B = csr_copy_in(Ne, Nf, B_rowptr, B_col, B_val);
    s->Bt = csr_transpose(&s->B);
    s->Wdiag = dup_d(Wdiag, (size_t)Ne);
    s->w_sqrt = dup_d(NULL, (size_t)Ne);
    s->negaW = dup_d(NULL, (size_t)Ne);
    for (int i = 0; i < Ne; ++i) {
        s->w_sqrt[i] = sqrt(Wdiag[i]);      /* PDESampler.cpp:248-254 */
        s->negaW[i] = -1.0 * alpha * Wdiag[i]; /* :256-258 */
    }
    s->hasP = P_rowptr != NULL;
    if (s->hasP) {
        s->P = csr_copy_in(Ne, P_cols, P_rowptr, P_col, P_val);
        s->Pt = csr_transpose(&s->P);
    }
    s->set = 1;
    return 0;
}

int po_set_field_transfer(po_problem *p, int level, int n_out, const int *T_rowptr, const int *T_col,
                          const double *T_val, const double *row_scale)
{
    if (level < 0 || level >= p->nlevels || !p->s[level].set) return -1;
    sampler_level_t *s = &p->s[level];
    s->T = csr_copy_in(n_out, s->Ne, T_rowptr, T_col, T_val);
    s->tscale = (double *)malloc(sizeof(double) * (size_t)n_out);
    for (int i = 0; i < n_out; ++i) s->tscale[i] = row_scale ? row_scale[i] : 1.0;
    s->n_out = n_out;
    s->hasT = 1;
    return 0;
}

int po_set_darcy_level(po_problem *p, int level, int Ne, int Nf,
                       const int *elem_ptr, const int *elem_dofs, const double *elem_mat,
                       const int *B_rowptr, const int *B_col, const double *B_val,
                       const int *ess_u, const double *ess_data, const double *rhs, const double *obs,
                       int Pp_cols, const int *Pp_rowptr, const int *Pp_col, const double *Pp_val)
{
    if (level < 0 || level >= p->nlevels) return -1;
    darcy_level_t *d = &p->d[level];
    int N = Ne + Nf;
    d->Ne = Ne; d->Nf = Nf;
    int ndofs = elem_ptr[Ne];
    d->elem_ptr = (int *)malloc(sizeof(int) * (size_t)(Ne + 1));
    memcpy(d->elem_ptr, elem_ptr, sizeof(int) * (size_t)(Ne + 1));
    d->elem_dofs = (int *)malloc(sizeof(int) * (size_t)ndofs);
    memcpy(d->elem_dofs, elem_dofs, sizeof(int) * (size_t)ndofs);
    d->elem_mat_ptr = (long *)malloc(sizeof(long) * (size_t)(Ne + 1));
    d->elem_mat_ptr[0] = 0;
    for (int e = 0; e < Ne; ++e) {
        long n = elem_ptr[e + 1] - elem_ptr[e];
        d->elem_mat_ptr[e + 1] = d->elem_mat_ptr[e] + n * n;
    }
    d->elem_mat = dup_d(elem_mat, (size_t)d->elem_mat_ptr[Ne]);
    /* pattern of M: union of element dof pairs, via COO -> CSR with sorted, merged columns */
    {
        long nent = d->elem_mat_ptr[Ne];
        int *cnt = (int *)calloc((size_t)(Nf + 1), sizeof(int));
        for (int e = 0; e < Ne; ++e) {
            int n = elem_ptr[e + 1] - elem_ptr[e];
            for (int a = 0; a < n; ++a) cnt[elem_dofs[elem_ptr[e] + a] + 1] += n;
        }
        for (int i = 0; i < Nf; ++i) cnt[i + 1] += cnt[i];
        int *tmpcol = (int *)malloc(sizeof(int) * (size_t)(nent ? nent : 1));
        int *pos = (int *)malloc(sizeof(int) * (size_t)(Nf ? Nf : 1));
        memcpy(pos, cnt, sizeof(int) * (size_t)Nf);
        for (int e = 0; e < Ne; ++e) {
            int n = elem_ptr[e + 1] - elem_ptr[e];
            const int *dof = elem_dofs + elem_ptr[e];
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) tmpcol[pos[dof[a]]++] = dof[b];
        }
        csr_t Mp;
        Mp.rows = Nf; Mp.cols = Nf;
        Mp.rowptr = (int *)calloc((size_t)(Nf + 1), sizeof(int));
        Mp.col = (int *)malloc(sizeof(int) * (size_t)(nent ? nent : 1));
        int nnz = 0;
        for (int i = 0; i < Nf; ++i) {
            int s0 = cnt[i], s1 = cnt[i + 1];
            qsort(tmpcol + s0, (size_t)(s1 - s0), sizeof(int), cmp_int);
            for (int q = s0; q < s1; ++q)
                if (q == s0 || tmpcol[q] != tmpcol[q - 1]) Mp.col[nnz++] = tmpcol[q];
            Mp.rowptr[i + 1] = nnz;
        }
        Mp.val = (double *)calloc((size_t)(nnz ? nnz : 1), sizeof(double));
        d->Mpat = Mp;
        d->scatter = (int *)malloc(sizeof(int) * (size_t)(nent ? nent : 1));
        for (int e = 0; e < Ne; ++e) {
            int n = elem_ptr[e + 1] - elem_ptr[e];
            const int *dof = elem_dofs + elem_ptr[e];
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) {
                    int i = dof[a], j = dof[b];
                    int lo = Mp.rowptr[i], hi = Mp.rowptr[i + 1] - 1;
                    while (lo < hi) {
                        int mid = (lo + hi) / 2;
                        if (Mp.col[mid] < j) lo = mid + 1; else hi = mid;
                    }
                    d->scatter[d->elem_mat_ptr[e] + (long)a * n + b] = lo;
                }
        }
        free(cnt); free(tmpcol); free(pos);
    }
    d->B = csr_copy_in(Ne, Nf, B_rowptr, B_col, B_val);
    d->Be = csr_copy_in(Ne, Nf, B_rowptr, B_col, B_val);
    d->ess_u = (int *)malloc(sizeof(int) * (size_t)Nf);
    memcpy(d->ess_u, ess_u, sizeof(int) * (size_t)Nf);
    for (int i = 0; i < Ne; ++i)
        for (int q = d->Be.rowptr[i]; q < d->Be.rowptr[i + 1]; ++q)
            if (ess_u[d->Be.col[q]]) d->Be.val[q] = 0.0;
    d->Bet = csr_transpose(&d->Be);
    d->ess_data = dup_d(ess_data, (size_t)N);
    d->rhs = dup_d(rhs, (size_t)N);
    d->obs = dup_d(obs, (size_t)N);
    d->hasP = Pp_rowptr != NULL;
    if (d->hasP) d->Pp = csr_copy_in(Ne, Pp_cols, Pp_rowptr, Pp_col, Pp_val);
    d->set = 1;
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
static void sampler_prepare(po_problem *p, int level)
{
    sampler_level_t *s = &p->s[level];
    if (s->mg_ready) return;
    double *Md = (double *)malloc(sizeof(double) * (size_t)s->Nf);
    for (int i = 0; i < s->Nf; ++i) Md[i] = csr_diag_entry(&s->M, i);
    double *aW = (double *)malloc(sizeof(double) * (size_t)s->Ne);
    for (int i = 0; i < s->Ne; ++i) aW[i] = s->alpha * s->Wdiag[i];
    csr_t S = schur_diag(&s->B, &s->Bt, Md, aW);
    const csr_t *P[PO_MAX_MG];
    int nP = 0;
    for (int m = level; m < p->nlevels - 1 && p->s[m].hasP && nP < PO_MAX_MG - 1; ++m) P[nP++] = &p->s[m].P;
    mg_setup(&s->mg, S, P, nP);
    s->mg_ready = 1;
    free(Md); free(aW);
}

int po_sampler_eval(po_problem *p, int level, int xi_level, const double *xi, double *s_out,
                    double *embed_s, int init_level, int use_init, int *iters)
{
    if (level < 0 || level >= p->nlevels || xi_level > level || xi_level < 0) return -1;
    sampler_level_t *sl = &p->s[level];
    if (!sl->set) return -2;
#ifdef _OPENMP
#pragma omp critical(po_sampler_prepare)
#endif
    sampler_prepare(p, level);
    int Ne = sl->Ne, Nf = sl->Nf, N = Ne + Nf;
    /* rhs_s = -g W^{1/2} xi   (PDESampler.cpp:352-358 / :423-428) */
    int n0 = p->s[xi_level].Ne;
    double *rhs_s = (double *)malloc(sizeof(double) * (size_t)n0);
    for (int i = 0; i < n0; ++i) rhs_s[i] = -p->s[xi_level].g * xi[i] * p->s[xi_level].w_sqrt[i];
    /* project to the correct level (:361-368 / :431-438) */
    for (int l = xi_level; l < level; ++l) {
        double *tmp = (double *)malloc(sizeof(double) * (size_t)p->s[l + 1].Ne);
        csr_mult_transpose(&p->s[l].P, rhs_s, tmp);
        free(rhs_s);
        rhs_s = tmp;
    }
    double *b = (double *)calloc((size_t)N, sizeof(double));
    double *x = (double *)calloc((size_t)N, sizeof(double));
    memcpy(b + Nf, rhs_s, sizeof(double) * (size_t)Ne);
    int iterative = 0;
    if (use_init > 0 && embed_s) {
        /* prolongate the coarser Gaussian field (:496-508) */
        int il = init_level;
        double *cur = dup_d(embed_s, (size_t)p->s[il].Ne);
        while (il > level) {
            double *fine = (double *)malloc(sizeof(double) * (size_t)p->s[il - 1].Ne);
            csr_mult(&p->s[il - 1].P, cur, fine);
            free(cur);
            cur = fine;
            --il;
        }
        memcpy(x + Nf, cur, sizeof(double) * (size_t)Ne);
        free(cur);
        iterative = 1;
    }
    /* the level's multigrid hierarchy is shared and read-only; its scratch vectors are per call so that
     * concurrent samples (OpenMP threads of po_mlmc_level) do not share work space */
    mg_t mgl = sl->mg;
    for (int m = 0; m < mgl.nlev; ++m) {
        size_t n = (size_t)mgl.S[m].rows;
        mgl.r[m] = (double *)malloc(sizeof(double) * n);
        mgl.x[m] = (double *)malloc(sizeof(double) * n);
        mgl.t[m] = (double *)malloc(sizeof(double) * n);
    }
    saddle_t A = {Nf, Ne, &sl->M, &sl->B, &sl->Bt, sl->negaW, &mgl, 3};
    int it = minres(&A, b, x, iterative, p->rel, p->abs_, p->maxit);
    for (int m = 0; m < mgl.nlev; ++m) { free(mgl.r[m]); free(mgl.x[m]); free(mgl.t[m]); }
    if (iters) *iters = it;
    if (embed_s) memcpy(embed_s, x + Nf, sizeof(double) * (size_t)Ne); /* :527 */
    if (sl->hasT) { /* project to the original mesh, then exp (EmbeddedPDESampler.cpp:426-435) */
        double *t = (double *)malloc(sizeof(double) * (size_t)sl->n_out);
        csr_mult(&sl->T, x + Nf, t);
        for (int i = 0; i < sl->n_out; ++i) {
            const double v = t[i] * sl->tscale[i];
            s_out[i] = sl->lognormal ? exp(v) : v;
        }
        free(t);
    } else
        for (int i = 0; i < Ne; ++i) s_out[i] = sl->lognormal ? exp(x[Nf + i]) : x[Nf + i]; /* :529-533 */
    free(rhs_s); free(b); free(x);
    return 0;
}

int po_darcy_solve(po_problem *p, int level, const double *k, double *Q, double *C, double *sol_out,
                   int *iters)
{
    if (level < 0 || level >= p->nlevels) return -1;
    darcy_level_t *d = &p->d[level];
    if (!d->set) return -2;
    int Ne = d->Ne, Nf = d->Nf, N = Ne + Nf;
    /* M = ComputeMassOperator(uform, k)  (DarcySolver.cpp:479) */
    csr_t M = csr_copy_in(Nf, Nf, d->Mpat.rowptr, d->Mpat.col, NULL);
    for (int e = 0; e < Ne; ++e) {
        long o = d->elem_mat_ptr[e], n2 = d->elem_mat_ptr[e + 1] - o;
        for (long q = 0; q < n2; ++q) M.val[d->scatter[o + q]] += k[e] * d->elem_mat[o + q];
    }
    /* rhs_bc = rhs; EliminateRowCol(ess, ess_data, rhs_bc)  (:495-498) */
    double *b = dup_d(d->rhs, (size_t)N);
    for (int i = 0; i < Nf; ++i)
        for (int q = M.rowptr[i]; q < M.rowptr[i + 1]; ++q) {
            int j = M.col[q];
            if (d->ess_u[j] && !d->ess_u[i]) b[i] -= M.val[q] * d->ess_data[j];
            if (d->ess_u[j] || d->ess_u[i]) M.val[q] = (i == j) ? 1.0 : 0.0;
        }
    for (int i = 0; i < Ne; ++i)
        for (int q = d->B.rowptr[i]; q < d->B.rowptr[i + 1]; ++q) {
            int j = d->B.col[q];
            if (d->ess_u[j]) b[Nf + i] -= d->B.val[q] * d->ess_data[j];
        }
    for (int j = 0; j < Nf; ++j)
        if (d->ess_u[j]) b[j] = d->ess_data[j];
    /* preconditioner (re)build per sample (:568-601) */
    double *Md = (double *)malloc(sizeof(double) * (size_t)Nf);
    for (int i = 0; i < Nf; ++i) Md[i] = csr_diag_entry(&M, i);
    csr_t S = schur_diag(&d->Be, &d->Bet, Md, NULL);
    const csr_t *P[PO_MAX_MG];
    int nP = 0;
    for (int m = level; m < p->nlevels - 1 && p->d[m].hasP && nP < PO_MAX_MG - 1; ++m) P[nP++] = &p->d[m].Pp;
    mg_t mg;
    mg_setup(&mg, S, P, nP);
    saddle_t A = {Nf, Ne, &M, &d->Be, &d->Bet, NULL, &mg, 3};
    double *x = (double *)calloc((size_t)N, sizeof(double));
    int it = minres(&A, b, x, 0, p->rel, p->abs_, p->maxit); /* :629-631 */
    if (iters) *iters = it;
    *Q = vdot(N, d->obs, x); /* :427 */
    *C = (double)N;          /* :429 */
    if (sol_out) memcpy(sol_out, x, sizeof(double) * (size_t)N);
    mg_free(&mg);
    csr_free(&M);
    free(Md); free(b); free(x);
    return 0;
}

int po_mlmc_level(po_problem *p, int level, int nlevels, int nsamples, uint64_t pos0,
                  double mu, double sigma, double *sums, double *rows, int nthreads,
                  int64_t *total_iters)
{
    if (level < 0 || level >= nlevels || nlevels > p->nlevels) return -1;
    int coarsest = (level == nlevels - 1);
    int Ne = p->s[level].Ne;
    double *loc = (double *)malloc(sizeof(double) * 4 * (size_t)(nsamples > 0 ? nsamples : 1));
    int64_t its = 0;
    /* make the lazily built sampler hierarchies before going parallel */
    sampler_prepare(p, level);
    if (!coarsest) sampler_prepare(p, level + 1);
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : its)
#endif
    for (int j = 0; j < nsamples; ++j) {
        po_yarn5 g;
        po_yarn5_init(&g);
        po_yarn5_jump(&g, pos0 + (uint64_t)j * (uint64_t)Ne);
        double *xi = (double *)malloc(sizeof(double) * (size_t)Ne);
        po_normal_fill(&g, mu, sigma, Ne, xi); /* sampler.Sample(ilevel, xi) */
        double q = 0, c = 0, qc = 0, cc = 0;
        int it;
        if (coarsest) { /* MLMC_Manager.cpp:113-136 */
            double *sp_ = (double *)malloc(sizeof(double) * (size_t)(Ne + p->d[level].Ne));
            po_sampler_eval(p, level, level, xi, sp_, NULL, 0, -1, &it); its += it;
            po_darcy_solve(p, level, sp_, &q, &c, NULL, &it); its += it;
            free(sp_);
            loc[4 * j + 0] = q; loc[4 * j + 1] = q; loc[4 * j + 2] = 0.0; loc[4 * j + 3] = c;
        } else { /* :144-173 */
            int Nec = p->s[level + 1].Ne;
            double *spc = (double *)malloc(sizeof(double) * (size_t)(Nec + p->d[level + 1].Ne));
            double *spf = (double *)malloc(sizeof(double) * (size_t)(Ne + p->d[level].Ne));
            double *init = (double *)malloc(sizeof(double) * (size_t)(Ne > Nec ? Ne : Nec));
            po_sampler_eval(p, level + 1, level, xi, spc, init, 0, 0, &it); its += it;
            po_darcy_solve(p, level + 1, spc, &qc, &cc, NULL, &it); its += it;
            po_sampler_eval(p, level, level, xi, spf, init, level + 1, 1, &it); its += it;
            po_darcy_solve(p, level, spf, &q, &c, NULL, &it); its += it;
            free(spc); free(spf); free(init);
            loc[4 * j + 0] = q - qc; loc[4 * j + 1] = q; loc[4 * j + 2] = qc; loc[4 * j + 3] = c + cc;
        }
        free(xi);
    }
    for (int j = 0; j < nsamples; ++j) {
        double y = loc[4 * j], q = loc[4 * j + 1], c = loc[4 * j + 3];
        sums[7] += y * y * y;      /* Y3 */
        sums[8] += y * y * y * y;  /* Y4 */
        sums[0] += y * y;          /* Y2 */
        sums[1] += y;              /* Y  */
        sums[2] += fabs(y);        /* ABSY */
        sums[3] += q * q;          /* Q2 */
        sums[4] += q;              /* Q  */
        sums[5] += fabs(q);        /* ABSQ */
        sums[6] += c;              /* C  */
    }
    if (rows) memcpy(rows, loc, sizeof(double) * 4 * (size_t)nsamples);
    if (total_iters) *total_iters = its;
    free(loc);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* BayesianInverseProblem + ML_BayesRatio_Manager                                                 */
int po_set_observations(po_problem *p, int level, int m, const double *g, const double *G_obs, double noise)
{
    if (level < 0 || level >= p->nlevels || !p->d[level].set || m < 1) return -1;
    darcy_level_t *d = &p->d[level];
    d->n_obs = m;
    d->gobs_func = dup_d(g, (size_t)m * (size_t)d->Ne);
    d->Gobs = dup_d(G_obs, (size_t)m);
    d->noise = noise;
    return 0;
}

/* ComputeLikelihoodAndQ (/root/reference/src/BayesianInverseProblem.cpp:178-210): G_i = g_i . p / sum(g_i),
 * likelihood = exp(-|G - G_obs|^2 / (2 noise)); Q = obs . sol */
static void bayes_likelihood(po_problem *p, int level, const double *k, double *like, double *Q, int *iters)
{
    darcy_level_t *d = &p->d[level];
    double C;
    double *sol = (double *)malloc(sizeof(double) * (size_t)(d->Ne + d->Nf));
    po_darcy_solve(p, level, k, Q, &C, sol, iters);
    double n2 = 0.0;
    for (int i = 0; i < d->n_obs; ++i) {
        const double *g = d->gobs_func + (size_t)i * d->Ne;
        double num = 0.0, den = 0.0;
        for (int e = 0; e < d->Ne; ++e) { num += g[e] * sol[d->Nf + e]; den += g[e]; }
        const double dl = num / den - d->Gobs[i];
        n2 += dl * dl;
    }
    *like = exp((-1. / (d->noise * 2)) * n2);
    free(sol);
}

int po_bayes_level(po_problem *p, int level, int nlevels, int nsamples, uint64_t pos0, double mu, double sigma,
                   double *sums, double *rows, int nthreads)
{
    if (level < 0 || level >= nlevels || nlevels > p->nlevels) return -1;
    const int coarsest = (level == nlevels - 1);
    const int Ne = p->s[level].Ne;
    double *loc = (double *)malloc(sizeof(double) * 5 * (size_t)(nsamples > 0 ? nsamples : 1));
    sampler_prepare(p, level);
    if (!coarsest) sampler_prepare(p, level + 1);
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int j = 0; j < nsamples; ++j) {
        po_yarn5 g;
        po_yarn5_init(&g);
        po_yarn5_jump(&g, pos0 + (uint64_t)j * 2 * (uint64_t)Ne);
        double *zxi = (double *)malloc(sizeof(double) * (size_t)Ne);
        double *xi = (double *)malloc(sizeof(double) * (size_t)Ne);
        po_normal_fill(&g, mu, sigma, Ne, zxi); /* problem.SamplePrior(ilevel, zxi)  (ML_BayesRatio_Manager.hpp:334) */
        po_normal_fill(&g, mu, sigma, Ne, xi);  /* problem.SamplePrior(ilevel, xi)   (:339) */
        int nk = p->d[level].Ne + Ne + (coarsest ? 0 : p->d[level + 1].Ne + p->s[level + 1].Ne);
        double *par = (double *)malloc(sizeof(double) * (size_t)nk);
        double z = 0, r = 0, zc = 0, rc = 0, q = 0, c_tot = 0;
        int it;
        po_sampler_eval(p, level, level, zxi, par, NULL, 0, -1, &it);   /* EvalPrior: 3-argument Eval */
        bayes_likelihood(p, level, par, &z, &q, &it); c_tot += p->d[level].Ne + p->d[level].Nf;
        po_sampler_eval(p, level, level, xi, par, NULL, 0, -1, &it);
        bayes_likelihood(p, level, par, &r, &q, &it); r *= q; c_tot += p->d[level].Ne + p->d[level].Nf;   /* ComputeR */
        if (!coarsest) {
            po_sampler_eval(p, level + 1, level, zxi, par, NULL, 0, -1, &it);
            bayes_likelihood(p, level + 1, par, &zc, &q, &it); c_tot += p->d[level + 1].Ne + p->d[level + 1].Nf;
            po_sampler_eval(p, level + 1, level, xi, par, NULL, 0, -1, &it);
            bayes_likelihood(p, level + 1, par, &rc, &q, &it); rc *= q; c_tot += p->d[level + 1].Ne + p->d[level + 1].Nf;
        }
        loc[5 * j + 0] = r; loc[5 * j + 1] = coarsest ? r : r - rc;
        loc[5 * j + 2] = z; loc[5 * j + 3] = coarsest ? z : z - zc; loc[5 * j + 4] = c_tot;
        free(zxi); free(xi); free(par);
    }
    /* enum {YZ2, YZ, ABS_YZ, Z2, Z, ABS_Z, YR2, YR, ABS_YR, R2, R, ABS_R, ..., C = 18} (ML_BayesRatio_Manager.hpp:67-70) */
    for (int j = 0; j < nsamples; ++j) {
        const double r = loc[5 * j], yr = loc[5 * j + 1], z = loc[5 * j + 2], yz = loc[5 * j + 3];
        sums[10] += r; sums[11] += fabs(r); sums[9] += r * r;
        sums[7] += yr; sums[8] += fabs(yr); sums[6] += yr * yr;
        sums[4] += z; sums[5] += fabs(z); sums[3] += z * z;
        sums[1] += yz; sums[2] += fabs(yz); sums[0] += yz * yz;
        sums[18] += loc[5 * j + 4];
    }
    if (rows) memcpy(rows, loc, sizeof(double) * 5 * (size_t)nsamples);
    free(loc);
    return 0;
}

/* ============================================================================================ */
/* manager statistics                                                                           */
/* ============================================================================================ */
double po_exp_w_regression(const double *y, const double *x, int size, int skip_n_last)
{
    int n = size - 1 - skip_n_last;
    if (n < 1) return 0.0;
    double num = 0.0, den = 0.0;
    for (int i = 0; i < n; ++i) {
        double logdy = log(fabs(y[i] / y[i + 1]));
        double logdx = log(x[i] / x[i + 1]);
        double w = pow(.5, i);
        num += logdy * w * logdx;
        den += logdx * w * logdx;
    }
    return num / den;
}

void po_mlmc_compute(int nlevels, const double *sums, const int *nsamples, const double *M,
                     const double *cost_in, double eps2_in, double ratio,
                     double *eY, double *eABSY, double *eQ, double *eABSQ, double *eC, double *varY,
                     double *varQ, double *consistency, double *kurtosis, double *VC, int *missing,
                     po_mlmc_stats *out)
{
    enum { Y2 = 0, Y = 1, ABSY = 2, Q2 = 3, Q = 4, ABSQ = 5, C = 6, Y3 = 7, Y4 = 8, NVAR = 9 };
    double eps2 = eps2_in;
    int auto_eps2 = eps2 < 0 ? 1 : 0;
    for (int l = 0; l < nlevels; ++l) {
        double n = (double)nsamples[l];
        const double *s = sums + (size_t)l * NVAR;
        eY[l] = s[Y] / n; eABSY[l] = s[ABSY] / n; eQ[l] = s[Q] / n; eABSQ[l] = s[ABSQ] / n;
        eC[l] = s[C] / n;
        varY[l] = s[Y2] / n; varQ[l] = s[Q2] / n; kurtosis[l] = s[Y4] / n;
        kurtosis[l] /= varY[l] * varY[l];
        varY[l] -= eY[l] * eY[l];
        varY[l] *= n / (double)(nsamples[l] - 1);
        varQ[l] -= eQ[l] * eQ[l];
        varQ[l] *= n / (double)(nsamples[l] - 1);
        consistency[l] = 0.0;
    }
    for (int l = 0; l < nlevels - 1; ++l)
        consistency[l] = fabs(eQ[l] - eQ[l + 1] + eY[l]) /
                         (3 * (sqrt(varQ[l]) + sqrt(varQ[l + 1]) + sqrt(varY[l])));
    out->alpha = po_exp_w_regression(eY, M, nlevels, 1);
    out->alpha_abs = po_exp_w_regression(eABSY, M, nlevels, 1);
    out->beta = po_exp_w_regression(varY, M, nlevels, 1);
    double bias2 = 0.0;
    if (nlevels > 1) {
        double m = M[0] / M[1];
        if (nlevels > 3)
            bias2 = fmax(pow(m, 2. * out->alpha_abs) * eABSY[1] * eABSY[1], eABSY[0] * eABSY[0]) /
                    (pow(pow(m, -2. * out->alpha_abs) - 1., 2));
        else if (nlevels == 3)
            bias2 = (eABSY[0] * eABSY[0]) / (pow(pow(m, -out->alpha_abs) - 1., 2));
        else
            bias2 = eABSY[0] * eABSY[0];
    }
    if (auto_eps2) eps2 = bias2 / (1. - ratio);
    double mlvar = 0.0, est = 0.0;
    for (int l = 0; l < nlevels; ++l) { mlvar += varY[l] / (double)nsamples[l]; est += eY[l]; }
    const double *cost = cost_in ? cost_in : eC;
    out->gamma = po_exp_w_regression(cost, M, nlevels, 0);
    double prop = 0.0;
    for (int l = 0; l < nlevels; ++l) prop += sqrt(varY[l] * cost[l]);
    prop /= ratio * eps2;
    for (int l = 0; l < nlevels; ++l) {
        double missings = prop * sqrt(varY[l] / cost[l]);
        missings -= (double)nsamples[l];
        int mi = (int)ceil(missings);
        missing[l] = mi > 0 ? mi : 0;
        VC[l] = varY[l] * cost[l];
    }
    out->estimate = est; out->ml_estimator_variance = mlvar; out->bias2 = bias2;
    out->actual_mse = bias2 + mlvar; out->eps2 = eps2;
}

void po_mc_compute(const double *sums, int nsamples, const double *cost_in, double eps2_in, double ratio,
                   double *eQ, double *eABSQ, double *eC, double *varQ, int *missing, po_mlmc_stats *out)
{
    enum { Q2 = 0, Q = 1, ABSQ = 2, C = 3 };
    const double nl = (double)nsamples;
    double eps2 = eps2_in;
    int auto_eps2 = eps2 < 0 ? 1 : 0;
    *eQ = sums[Q] / nl; *eABSQ = sums[ABSQ] / nl; *eC = sums[C] / nl; *varQ = sums[Q2] / nl;
    *varQ -= (*eQ) * (*eQ);
    *varQ *= nl / (nl - 1.);
    double bias2 = 0.0;
    if (auto_eps2) eps2 = bias2 / (1. - ratio);
    double var = *varQ / nl;
    double cost = cost_in ? *cost_in : *eC;
    const double prop = sqrt(*varQ * cost) / (ratio * eps2);
    const double missings = prop * sqrt(*varQ / cost) - nl;
    int mi = (int)ceil(missings);
    *missing = mi > 0 ? mi : 0;
    memset(out, 0, sizeof *out);
    out->estimate = *eQ; out->ml_estimator_variance = var; out->bias2 = bias2;
    out->actual_mse = bias2 + var; out->eps2 = eps2;
}
