// host_sparse.hpp -- small host-side CSR toolkit used once per upload to derive the device operators
// (block saddle matrices, Schur-complement hierarchies, value maps) from the arrays ParELAG hands over.
// Nothing here runs per sample.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace pmc {

struct HCsr {
    int rows = 0, cols = 0;
    std::vector<int> rowptr{0};
    std::vector<int> col;
    std::vector<double> val;
    int nnz() const { return (int)col.size(); }
};

struct Coo {
    int r, c;
    double v;
};

// COO -> CSR, columns sorted within rows, duplicates summed.
inline HCsr csr_from_coo(int rows, int cols, std::vector<Coo> &e, bool drop_zeros = false)
{
    std::sort(e.begin(), e.end(), [](const Coo &a, const Coo &b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
    HCsr A;
    A.rows = rows;
    A.cols = cols;
    A.rowptr.assign(rows + 1, 0);
    size_t i = 0;
    while (i < e.size()) {
        size_t j = i;
        double s = 0;
        while (j < e.size() && e[j].r == e[i].r && e[j].c == e[i].c) s += e[j++].v;
        if (!(drop_zeros && s == 0.0)) {
            A.col.push_back(e[i].c);
            A.val.push_back(s);
            A.rowptr[e[i].r + 1]++;
        }
        i = j;
    }
    for (int r = 0; r < rows; ++r) A.rowptr[r + 1] += A.rowptr[r];
    return A;
}

inline HCsr csr_copy(int rows, int cols, const int *rp, const int *ci, const double *v)
{
    HCsr A;
    A.rows = rows;
    A.cols = cols;
    A.rowptr.assign(rp, rp + rows + 1);
    A.col.assign(ci, ci + rp[rows]);
    A.val.assign(v, v + rp[rows]);
    return A;
}

inline HCsr csr_transpose(const HCsr &A)
{
    HCsr T;
    T.rows = A.cols;
    T.cols = A.rows;
    T.rowptr.assign(T.rows + 1, 0);
    for (int c : A.col) T.rowptr[c + 1]++;
    for (int i = 0; i < T.rows; ++i) T.rowptr[i + 1] += T.rowptr[i];
    T.col.resize(A.nnz());
    T.val.resize(A.nnz());
    std::vector<int> pos(T.rowptr.begin(), T.rowptr.end() - 1);
    for (int i = 0; i < A.rows; ++i)
        for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
            int q = pos[A.col[p]]++;
            T.col[q] = i;
            T.val[q] = A.val[p];
        }
    return T;
}

// C = A * B, sorted columns.
inline HCsr csr_matmul(const HCsr &A, const HCsr &B)
{
    HCsr C;
    C.rows = A.rows;
    C.cols = B.cols;
    C.rowptr.assign(C.rows + 1, 0);
    std::vector<int> mark(C.cols, -1);
    std::vector<double> acc(C.cols, 0.0);
    std::vector<int> cols_here;
    for (int i = 0; i < A.rows; ++i) {
        cols_here.clear();
        for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
            int k = A.col[p];
            double a = A.val[p];
            for (int q = B.rowptr[k]; q < B.rowptr[k + 1]; ++q) {
                int j = B.col[q];
                if (mark[j] != i) {
                    mark[j] = i;
                    acc[j] = a * B.val[q];
                    cols_here.push_back(j);
                } else
                    acc[j] += a * B.val[q];
            }
        }
        std::sort(cols_here.begin(), cols_here.end());
        for (int j : cols_here) {
            C.col.push_back(j);
            C.val.push_back(acc[j]);
        }
        C.rowptr[i + 1] = (int)C.col.size();
    }
    return C;
}

inline std::vector<double> csr_diag(const HCsr &A)
{
    std::vector<double> d(A.rows, 0.0);
    for (int i = 0; i < A.rows; ++i)
        for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p)
            if (A.col[p] == i) d[i] += A.val[p];
    return d;
}

inline void csr_mult(const HCsr &A, const double *x, double *y)
{
    for (int i = 0; i < A.rows; ++i) {
        double s = 0;
        for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) s += A.val[p] * x[A.col[p]];
        y[i] = s;
    }
}

// Largest eigenvalue of D^-1 A (A symmetric positive semi-definite, D = diag d > 0) by power iteration.
inline double lambda_max_scaled(const HCsr &A, const std::vector<double> &d, int iters = 60)
{
    int n = A.rows;
    if (n == 0) return 1.0;
    std::vector<double> x(n), y(n);
    uint64_t s = 0x9E3779B97F4A7C15ULL;
    for (int i = 0; i < n; ++i) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        x[i] = 0.5 + (double)(s >> 40) / (double)(1 << 24);
    }
    double lam = 1.0;
    for (int it = 0; it < iters; ++it) {
        csr_mult(A, x.data(), y.data());
        double nrm = 0;
        for (int i = 0; i < n; ++i) {
            y[i] /= d[i];
            nrm = std::max(nrm, std::fabs(y[i]));
        }
        if (nrm == 0) return 1.0;
        lam = nrm;
        for (int i = 0; i < n; ++i) x[i] = y[i] / nrm;
    }
    // Rayleigh-type refinement in the D inner product
    csr_mult(A, x.data(), y.data());
    double num = 0, den = 0;
    for (int i = 0; i < n; ++i) {
        num += x[i] * y[i];
        den += x[i] * d[i] * x[i];
    }
    if (den > 0) lam = std::max(lam, num / den);
    return lam;
}

// Largest eigenvalue of diag(A)^-1 A for a small dense symmetric positive (semi-)definite n x n block.
inline double dense_lambda_max_scaled(const double *A, int n, int iters = 200)
{
    std::vector<double> x(n, 1.0), y(n);
    for (int i = 0; i < n; ++i) x[i] = 1.0 + 0.37 * ((i * 7) % 5);
    double lam = 1.0;
    for (int it = 0; it < iters; ++it) {
        double nrm = 0;
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int j = 0; j < n; ++j) s += A[i * n + j] * x[j];
            double d = A[i * n + i];
            y[i] = d > 0 ? s / d : 0.0;
            nrm = std::max(nrm, std::fabs(y[i]));
        }
        if (nrm == 0) return 1.0;
        lam = nrm;
        for (int i = 0; i < n; ++i) x[i] = y[i] / nrm;
    }
    // Gershgorin-type cap keeps this an upper bound even if the power iteration has not converged
    double cap = 0;
    for (int i = 0; i < n; ++i) {
        double s = 0, d = A[i * n + i];
        for (int j = 0; j < n; ++j) s += std::fabs(A[i * n + j]);
        if (d > 0) cap = std::max(cap, s / d);
    }
    return std::max(1.0, std::min(1.05 * lam, cap));
}

// Extreme eigenvalues of a symmetric positive definite operator given by its action (Lanczos without
// reorthogonalisation: the extreme Ritz values converge first and ghost copies do not move them; the extreme
// eigenvalues of the Lanczos tridiagonal matrix by Sturm-sequence bisection).  lo_out >= lambda_min and hi_out <=
// lambda_max (Ritz values lie inside the spectrum): callers widen the interval by a safety margin.
template <typename Apply>
inline void lanczos_extremes(int n, int steps, Apply &&apply, double *lo_out, double *hi_out)
{
    *lo_out = *hi_out = 1.0;
    if (n <= 0) return;
    steps = std::max(1, std::min(steps, n));
    std::vector<double> v(n), vp(n, 0.0), w(n), al, be;
    uint64_t s = 0x9E3779B97F4A7C15ULL;
    double nrm = 0;
    for (int i = 0; i < n; ++i) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        v[i] = (double)(s >> 40) / (double)(1 << 24) - 0.5;
        nrm += v[i] * v[i];
    }
    nrm = std::sqrt(nrm);
    for (int i = 0; i < n; ++i) v[i] /= nrm;
    double beta = 0;
    for (int k = 0; k < steps; ++k) {
        apply(v.data(), w.data());
        double a = 0;
        for (int i = 0; i < n; ++i) a += w[i] * v[i];
        double b2 = 0;
        for (int i = 0; i < n; ++i) {
            w[i] -= a * v[i] + beta * vp[i];
            b2 += w[i] * w[i];
        }
        al.push_back(a);
        beta = std::sqrt(b2);
        if (beta <= 1e-14 * std::fabs(a) || k + 1 == steps) break;
        be.push_back(beta);
        for (int i = 0; i < n; ++i) {
            vp[i] = v[i];
            v[i] = w[i] / beta;
        }
    }
    const int m = (int)al.size();
    double glo = al[0], ghi = al[0];   // Gershgorin interval of the tridiagonal matrix
    for (int i = 0; i < m; ++i) {
        const double r = (i > 0 ? std::fabs(be[i - 1]) : 0.0) + (i + 1 < m ? std::fabs(be[i]) : 0.0);
        glo = std::min(glo, al[i] - r);
        ghi = std::max(ghi, al[i] + r);
    }
    auto count_below = [&](double x) {   // number of eigenvalues of T below x
        int c = 0;
        double q = al[0] - x;
        if (q < 0) ++c;
        for (int i = 1; i < m; ++i) {
            if (q == 0.0) q = 1e-300;
            q = al[i] - x - be[i - 1] * be[i - 1] / q;
            if (q < 0) ++c;
        }
        return c;
    };
    auto kth = [&](int k) {   // k-th smallest eigenvalue of T (0-based)
        double a = glo, b = ghi;
        for (int it = 0; it < 200 && b - a > 1e-14 * std::max(std::fabs(a), std::fabs(b)); ++it) {
            const double mid = 0.5 * (a + b);
            if (count_below(mid) > k) b = mid; else a = mid;
        }
        return 0.5 * (a + b);
    };
    *lo_out = kth(0);
    *hi_out = kth(m - 1);
}

}  // namespace pmc
