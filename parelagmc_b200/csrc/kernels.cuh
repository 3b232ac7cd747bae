// kernels.cuh -- device kernels of the batched per-sample path (sm_100a).
//
// DATA LAYOUT.  Every per-sample vector of a batch lives in ONE array X[row * ld + s]: `row` is the dof
// (RT face / L2 element / Schur value id), `s` the realisation, `ld` the (even) padded batch size.  The
// realisation index is the fastest-varying one, so a warp that works on one matrix row reads and writes
// 64 consecutive realisations = 512 contiguous bytes per access (one 128-bit load per lane), while the
// matrix entry (column index, coefficient) is the same for the whole warp and is fetched once through the
// read-only path.  All kernels are therefore bound by the streaming of the batched vectors through HBM;
// the sparse matrices themselves are amortised over the batch and stay in L2.
//
// THREAD MAPPING (all batched kernels).  blockDim = (TX, TY).  Thread (tx, ty) owns the realisation PAIR
// sp = blockIdx.x * TX + tx (realisations 2sp, 2sp+1, one double2).  The CTA owns the row block
// [blockIdx.y * rows_per_cta, ...) and its TY row lanes stride over it.  Dot products are reduced over the
// TY lanes in shared memory in a fixed order and written as per-row-block partial sums
// partial[(off + blockIdx.y) * ld + s]; a per-sample scalar kernel adds the partials in block order, so
// every reduction is deterministic and independent of how realisations are grouped into batches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmc {

constexpr int TX = 32;  // realisation pairs per CTA  (64 realisations)
constexpr int TY = 8;   // row lanes per CTA

__device__ __forceinline__ double2 ld2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ double2 ld2w(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

// Fixed-order reduction of one double2 per thread over the TY row lanes; lane ty == 0 returns the sum.
__device__ __forceinline__ double2 reduce_rows(double2 acc, double2 (*red)[TX])
{
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    double2 s = make_double2(0.0, 0.0);
    if (threadIdx.y == 0) {
#pragma unroll
        for (int j = 0; j < TY; ++j) {
            s.x += red[j][threadIdx.x].x;
            s.y += red[j][threadIdx.x].y;
        }
    }
    return s;
}

// ------------------------------------------------------------------------------------------------------
// Multi-RHS sparse operator apply with fused epilogues.
//   plain CSR   : row i has entries p in [rowptr[i], rowptr[i+1]):            sum += val[p] * x[col[p]]
//   weighted CSR: row i has WEIGHTED entries p in [rowptr[2i], rowptr[2i+1]): sum += val[p] * V[widx[p]] * x[col[p]]
//                 and FIXED entries   p in [rowptr[2i+1], rowptr[2i+2]):      sum += val[p] * x[col[p]]
//   V is a batched vector of per-sample weights: the permeability k_e for the Raviart-Thomas mass block
//   M(k) = sum_e k_e R_e^T M_e R_e (the batched element reassembly of DarcySolver::assemble,
//   /root/reference/src/DarcySolver.cpp:479, never materialised), or the per-sample Schur-complement values.
// ------------------------------------------------------------------------------------------------------
enum { EP_AX = 0, EP_RESID = 1, EP_CHEB = 2, EP_ADD = 3 };

struct SpmmArgs {
    int n, ld, rows_per_cta;
    const int *rowptr;
    const int *col;
    const double *val;
    const int *widx;
    const double *V;
    const double *x;     // gather source [cols][ld]
    double *y;           // EP_AX/EP_RESID: y ; EP_ADD: y += ; EP_CHEB: z_out
    const double *r;     // EP_RESID, EP_CHEB
    double *d;           // EP_CHEB direction (in/out)
    const double *dinv;  // EP_CHEB: [n] (fixed) or [n][ld] (batched)
    double ca, cb;       // EP_CHEB: d = ca d + cb dinv (r - A x);  z_out = x + d.   EP_ADD: y += ca (A x)
    const double *dotw;  // DOT: partial += out_row * dotw_row
    double *partial;
    int partial_off;
    const int *perm;     // optional processing order: the CTA's i-th row is perm[i] (storage order is unchanged)
};

template <int EP, bool WEIGHTED, bool BDINV, bool DOT>
__global__ void __launch_bounds__(TX *TY) k_spmm(const SpmmArgs a)
{
    __shared__ double2 red[DOT ? TY : 1][TX];
    const int sp = blockIdx.x * TX + threadIdx.x;
    const bool live = 2 * sp < a.ld;
    const size_t so = 2 * (size_t)sp;
    const size_t ld = (size_t)a.ld;
    const int r0 = blockIdx.y * a.rows_per_cta;
    const int r1 = min(a.n, r0 + a.rows_per_cta);
    double2 acc = make_double2(0.0, 0.0);
    if (live) {
        for (int ri = r0 + threadIdx.y; ri < r1; ri += TY) {
            const int row = a.perm ? __ldg(a.perm + ri) : ri;
            double2 s = make_double2(0.0, 0.0);
            if (WEIGHTED) {
                const int p0 = __ldg(a.rowptr + 2 * row), p1 = __ldg(a.rowptr + 2 * row + 1),
                          p2 = __ldg(a.rowptr + 2 * row + 2);
#pragma unroll 4
                for (int p = p0; p < p1; ++p) {
                    const double c = __ldg(a.val + p);
                    const double2 wv = ld2(a.V + (size_t)__ldg(a.widx + p) * ld + so);
                    const double2 xv = ld2(a.x + (size_t)__ldg(a.col + p) * ld + so);
                    s.x = fma(c * wv.x, xv.x, s.x);
                    s.y = fma(c * wv.y, xv.y, s.y);
                }
#pragma unroll 4
                for (int p = p1; p < p2; ++p) {
                    const double c = __ldg(a.val + p);
                    const double2 xv = ld2(a.x + (size_t)__ldg(a.col + p) * ld + so);
                    s.x = fma(c, xv.x, s.x);
                    s.y = fma(c, xv.y, s.y);
                }
            } else {
                const int p0 = __ldg(a.rowptr + row), p1 = __ldg(a.rowptr + row + 1);
#pragma unroll 4
                for (int p = p0; p < p1; ++p) {
                    const double c = __ldg(a.val + p);
                    const double2 xv = ld2(a.x + (size_t)__ldg(a.col + p) * ld + so);
                    s.x = fma(c, xv.x, s.x);
                    s.y = fma(c, xv.y, s.y);
                }
            }
            const size_t o = (size_t)row * ld + so;
            double2 out;
            if (EP == EP_AX) {
                out = s;
            } else if (EP == EP_RESID) {
                const double2 rv = ld2(a.r + o);
                out = make_double2(rv.x - s.x, rv.y - s.y);
            } else if (EP == EP_ADD) {
                const double2 yv = ld2w(a.y + o);
                out = make_double2(fma(a.ca, s.x, yv.x), fma(a.ca, s.y, yv.y));  // y += ca * (A x)
            } else {  // EP_CHEB
                const double2 rv = ld2(a.r + o);
                double2 di;
                if (BDINV) di = ld2(a.dinv + o);
                else { const double t = __ldg(a.dinv + row); di = make_double2(t, t); }
                double2 dn = make_double2(a.cb * di.x * (rv.x - s.x), a.cb * di.y * (rv.y - s.y));
                if (a.ca != 0.0) {
                    const double2 dv = ld2w(a.d + o);
                    dn.x = fma(a.ca, dv.x, dn.x);
                    dn.y = fma(a.ca, dv.y, dn.y);
                }
                st2(a.d + o, dn);
                const double2 zv = ld2(a.x + o);
                out = make_double2(zv.x + dn.x, zv.y + dn.y);
            }
            st2(a.y + o, out);
            if (DOT) {
                const double2 wv = ld2(a.dotw + o);
                acc.x = fma(out.x, wv.x, acc.x);
                acc.y = fma(out.y, wv.y, acc.y);
            }
        }
    }
    if (DOT) {
        const double2 s = reduce_rows(acc, red);
        if (threadIdx.y == 0 && live) st2(a.partial + (size_t)(a.partial_off + blockIdx.y) * ld + so, s);
    }
}

// First Chebyshev step from a zero initial guess: d = cb dinv r ; z = d  (no operator apply).
template <bool BDINV, bool DOT>
__global__ void __launch_bounds__(TX *TY)
    k_cheb_first(int n, int ld_, int rows_per_cta, const double *__restrict__ r, const double *__restrict__ dinv,
                 double cb, double *__restrict__ d, double *__restrict__ z, double *partial, int partial_off)
{
    __shared__ double2 red[DOT ? TY : 1][TX];
    const int sp = blockIdx.x * TX + threadIdx.x;
    const bool live = 2 * sp < ld_;
    const size_t so = 2 * (size_t)sp, ld = (size_t)ld_;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    double2 acc = make_double2(0.0, 0.0);
    if (live)
        for (int row = r0 + threadIdx.y; row < r1; row += TY) {
            const size_t o = (size_t)row * ld + so;
            const double2 rv = ld2(r + o);
            double2 di;
            if (BDINV) di = ld2(dinv + o);
            else { const double t = __ldg(dinv + row); di = make_double2(t, t); }
            const double2 dn = make_double2(cb * di.x * rv.x, cb * di.y * rv.y);
            st2(d + o, dn);
            st2(z + o, dn);
            if (DOT) {
                acc.x = fma(dn.x, rv.x, acc.x);
                acc.y = fma(dn.y, rv.y, acc.y);
            }
        }
    if (DOT) {
        const double2 s = reduce_rows(acc, red);
        if (threadIdx.y == 0 && live) st2(partial + (size_t)(partial_off + blockIdx.y) * ld + so, s);
    }
}

// Lanczos vector update of MINRES:  C = cA[s] A + cB[s] B + cC[s] C   (per-sample coefficients).
__global__ void __launch_bounds__(TX *TY)
    k_lincomb3(int n, int ld_, int rows_per_cta, const double *__restrict__ cA, const double *__restrict__ A,
               const double *__restrict__ cB, const double *__restrict__ B, const double *__restrict__ cC,
               double *__restrict__ Cv)
{
    const int sp = blockIdx.x * TX + threadIdx.x;
    if (2 * sp >= ld_) return;
    const size_t so = 2 * (size_t)sp, ld = (size_t)ld_;
    const double2 a = ld2(cA + so), b = ld2(cB + so), c = ld2(cC + so);
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
#pragma unroll 2
    for (int row = r0 + threadIdx.y; row < r1; row += TY) {
        const size_t o = (size_t)row * ld + so;
        const double2 av = ld2(A + o), bv = ld2(B + o);
        double2 cv = make_double2(0.0, 0.0);
        if (c.x != 0.0 || c.y != 0.0) cv = ld2w(Cv + o);
        double2 out;
        out.x = fma(a.x, av.x, fma(b.x, bv.x, c.x * cv.x));
        out.y = fma(a.y, av.y, fma(b.y, bv.y, c.y * cv.y));
        st2(Cv + o, out);
    }
}

// Search-direction and solution update of MINRES:
//   w0 = cw0[s] w0 + cw1[s] w1 + cu[s] u1 ;  x += cx[s] w0.
__global__ void __launch_bounds__(TX *TY)
    k_solution_update(int n, int ld_, int rows_per_cta, const double *__restrict__ cw0, const double *__restrict__ cw1,
                      const double *__restrict__ cu, const double *__restrict__ cx, double *__restrict__ w0,
                      const double *__restrict__ w1, const double *__restrict__ u1, double *__restrict__ x)
{
    const int sp = blockIdx.x * TX + threadIdx.x;
    if (2 * sp >= ld_) return;
    const size_t so = 2 * (size_t)sp, ld = (size_t)ld_;
    const double2 a = ld2(cw0 + so), b = ld2(cw1 + so), c = ld2(cu + so), e = ld2(cx + so);
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
#pragma unroll 2
    for (int row = r0 + threadIdx.y; row < r1; row += TY) {
        const size_t o = (size_t)row * ld + so;
        const double2 w0v = ld2w(w0 + o), w1v = ld2(w1 + o), uv = ld2(u1 + o);
        double2 xv = ld2w(x + o);
        double2 wn;
        wn.x = fma(a.x, w0v.x, fma(b.x, w1v.x, c.x * uv.x));
        wn.y = fma(a.y, w0v.y, fma(b.y, w1v.y, c.y * uv.y));
        st2(w0 + o, wn);
        if (e.x != 0.0) xv.x = fma(e.x, wn.x, xv.x);
        if (e.y != 0.0) xv.y = fma(e.y, wn.y, xv.y);
        st2(x + o, xv);
    }
}

// ------------------------------------------------------------------------------------------------------
// Per-sample scalar recurrences of preconditioned MINRES (mfem::MINRESSolver::Mult structure, as used through
// ParELAG's Krylov wrapper at /root/reference/src/PDESampler.cpp:517-522 and src/DarcySolver.cpp:629-631),
// with lazily normalised Lanczos vectors.  One thread per realisation.
// ------------------------------------------------------------------------------------------------------
enum {
    ST_BETA = 0, ST_IB, ST_IBPREV, ST_G0, ST_G1, ST_S0, ST_S1, ST_ETA, ST_GOAL, ST_ALPHA,
    ST_CQ, ST_CV1, ST_CV0, ST_CW0, ST_CW1, ST_CU, ST_CX, ST_COUNT
};

__device__ __forceinline__ double safe_inv(double x) { return x != 0.0 ? 1.0 / x : 0.0; }

// Sum of the row-block partials of one realisation, in block order (deterministic); loads are issued eight at a
// time so the loop is not bound by one memory latency per term.
__device__ __forceinline__ double sum_partials(const double *partial, int nblk, int ld, int s)
{
    double d = 0.0;
    int b = 0;
    for (; b + 8 <= nblk; b += 8) {
        double t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = partial[(size_t)(b + k) * ld + s];
#pragma unroll
        for (int k = 0; k < 8; ++k) d += t[k];
    }
    for (; b < nblk; ++b) d += partial[(size_t)b * ld + s];
    return d;
}

__global__ void k_minres_init(int ld, int nsamples, int nblk, const double *__restrict__ partial, double rel,
                              double abs_, double *__restrict__ st, int *__restrict__ active, int *__restrict__ iters,
                              int *__restrict__ n_active)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ld) return;
    const double dsum = sum_partials(partial, nblk, ld, s);
    const double eta = sqrt(fmax(dsum, 0.0));
    const double goal = fmax(rel * eta, abs_);
    const int act = (s < nsamples) && (eta > goal);
    st[ST_BETA * ld + s] = eta;
    st[ST_IB * ld + s] = safe_inv(eta);
    st[ST_IBPREV * ld + s] = 0.0;
    st[ST_G0 * ld + s] = 1.0;
    st[ST_G1 * ld + s] = 1.0;
    st[ST_S0 * ld + s] = 0.0;
    st[ST_S1 * ld + s] = 0.0;
    st[ST_ETA * ld + s] = eta;
    st[ST_GOAL * ld + s] = goal;
    active[s] = act;
    iters[s] = 0;
    if (act) atomicAdd(n_active, 1);
}

__global__ void k_minres_alpha(int ld, int nblk, const double *__restrict__ partial, double *__restrict__ st,
                               const int *__restrict__ active, int *__restrict__ n_active)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) *n_active = 0;  // re-counted by k_minres_beta of this iteration
    if (s >= ld) return;
    double cq = 0.0, cv1 = 0.0, cv0 = 0.0;
    if (active[s]) {
        const double dsum = sum_partials(partial, nblk, ld, s);
        const double ib = st[ST_IB * ld + s];
        const double alpha = dsum * ib * ib;
        st[ST_ALPHA * ld + s] = alpha;
        cq = ib;
        cv1 = -alpha * ib;
        cv0 = -st[ST_BETA * ld + s] * st[ST_IBPREV * ld + s];
    }
    st[ST_CQ * ld + s] = cq;
    st[ST_CV1 * ld + s] = cv1;
    st[ST_CV0 * ld + s] = cv0;
}

__global__ void k_minres_beta(int ld, int nblk, const double *__restrict__ partial, int max_iter,
                              double *__restrict__ st, int *__restrict__ active, int *__restrict__ iters,
                              int *__restrict__ n_active)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ld) return;
    double cw0 = 0.0, cw1 = 0.0, cu = 0.0, cx = 0.0;
    if (active[s]) {
        const double dsum = sum_partials(partial, nblk, ld, s);
        const double beta_new = sqrt(fmax(dsum, 0.0));
        const double beta = st[ST_BETA * ld + s], alpha = st[ST_ALPHA * ld + s], ib = st[ST_IB * ld + s];
        double g0 = st[ST_G0 * ld + s], g1 = st[ST_G1 * ld + s], s0 = st[ST_S0 * ld + s], s1 = st[ST_S1 * ld + s];
        double eta = st[ST_ETA * ld + s];
        const double delta = g1 * alpha - g0 * s1 * beta;
        const double rho3 = s0 * beta;
        const double rho2 = s1 * alpha + g0 * g1 * beta;
        const double rho1 = hypot(delta, beta_new);
        const double ir = safe_inv(rho1);
        cw0 = -rho3 * ir;
        cw1 = -rho2 * ir;
        cu = ib * ir;
        g0 = g1;
        g1 = delta * ir;
        cx = g1 * eta;
        s0 = s1;
        s1 = beta_new * ir;
        eta = -s1 * eta;
        const int it = iters[s] + 1;
        iters[s] = it;
        st[ST_G0 * ld + s] = g0;
        st[ST_G1 * ld + s] = g1;
        st[ST_S0 * ld + s] = s0;
        st[ST_S1 * ld + s] = s1;
        st[ST_ETA * ld + s] = eta;
        st[ST_IBPREV * ld + s] = ib;
        st[ST_BETA * ld + s] = beta_new;
        st[ST_IB * ld + s] = safe_inv(beta_new);
        if (fabs(eta) <= st[ST_GOAL * ld + s] || it >= max_iter || beta_new == 0.0) active[s] = 0;
        else atomicAdd(n_active, 1);
    }
    st[ST_CW0 * ld + s] = cw0;
    st[ST_CW1 * ld + s] = cw1;
    st[ST_CU * ld + s] = cu;
    st[ST_CX * ld + s] = cx;
}

// ------------------------------------------------------------------------------------------------------
// Set-up SpMM (once per solve, not per iteration): y = f(A g(x)) with optional |x| and reciprocal output.
// Used for diag M(k) = Dm k, Schur values V_0 = T_0 diag(M(k))^-1, V_{m+1} = T_{m+1} V_m, l1 row norms.
// ------------------------------------------------------------------------------------------------------
template <bool ABSX, bool RECIP>
__global__ void __launch_bounds__(TX *TY)
    k_spmm_setup(int n, int ld_, int rows_per_cta, const int *__restrict__ rowptr, const int *__restrict__ col,
                 const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
    const int sp = blockIdx.x * TX + threadIdx.x;
    if (2 * sp >= ld_) return;
    const size_t so = 2 * (size_t)sp, ld = (size_t)ld_;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    for (int row = r0 + threadIdx.y; row < r1; row += TY) {
        double2 s = make_double2(0.0, 0.0);
        const int p0 = __ldg(rowptr + row), p1 = __ldg(rowptr + row + 1);
#pragma unroll 4
        for (int p = p0; p < p1; ++p) {
            const double c = __ldg(val + p);
            double2 xv = ld2(x + (size_t)__ldg(col + p) * ld + so);
            if (ABSX) { xv.x = fabs(xv.x); xv.y = fabs(xv.y); }
            s.x = fma(c, xv.x, s.x);
            s.y = fma(c, xv.y, s.y);
        }
        if (RECIP) { s.x = safe_inv(s.x); s.y = safe_inv(s.y); }
        st2(y + (size_t)row * ld + so, s);
    }
}

// ------------------------------------------------------------------------------------------------------
// Misc batched kernels
// ------------------------------------------------------------------------------------------------------
__global__ void k_fill(double *p, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// dst[row][s] = s < nsamples ? vec[row] : 0
__global__ void __launch_bounds__(TX *TY)
    k_broadcast(int n, int ld_, int rows_per_cta, int nsamples, const double *__restrict__ vec, double *__restrict__ dst)
{
    const int sp = blockIdx.x * TX + threadIdx.x;
    if (2 * sp >= ld_) return;
    const size_t so = 2 * (size_t)sp, ld = (size_t)ld_;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    for (int row = r0 + threadIdx.y; row < r1; row += TY) {
        const double v = __ldg(vec + row);
        st2(dst + (size_t)row * ld + so, make_double2(2 * sp < nsamples ? v : 0.0, 2 * sp + 1 < nsamples ? v : 0.0));
    }
}

// dst (batched rows [dst_row0, dst_row0+n)) = op(src rows [src_row0, ...)):  MODE 0 copy, 1 exp.
template <int MODE>
__global__ void __launch_bounds__(TX *TY)
    k_map_rows(int n, int ld_, int rows_per_cta, const double *__restrict__ src, double *__restrict__ dst)
{
    const int sp = blockIdx.x * TX + threadIdx.x;
    if (2 * sp >= ld_) return;
    const size_t so = 2 * (size_t)sp, ld = (size_t)ld_;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    for (int row = r0 + threadIdx.y; row < r1; row += TY) {
        double2 v = ld2(src + (size_t)row * ld + so);
        if (MODE == 1) { v.x = exp(v.x); v.y = exp(v.y); }
        st2(dst + (size_t)row * ld + so, v);
    }
}

// Host layout [nsamples][n] (sample-major)  ->  batched [n][ld];  MODE 1 applies the SPDE right-hand-side
// scaling of PDESampler::Eval (/root/reference/src/PDESampler.cpp:352-358): out = (-g * xi) * w_sqrt[row].
template <int MODE>
__global__ void k_transpose_in(int n, int ld, int nsamples, const double *__restrict__ src, double *__restrict__ dst,
                               double neg_g, const double *__restrict__ w_sqrt)
{
    __shared__ double tile[32][33];
    const int row0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int s = s0 + j, row = row0 + threadIdx.x;
        tile[j][threadIdx.x] = (s < nsamples && row < n) ? src[(size_t)s * n + row] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int row = row0 + j, s = s0 + threadIdx.x;
        if (row < n && s < ld) {
            double v = tile[threadIdx.x][j];
            if (MODE == 1) v = __dmul_rn(__dmul_rn(neg_g, v), __ldg(w_sqrt + row));
            dst[(size_t)row * ld + s] = v;
        }
    }
}

// batched [n][ld] -> host layout [nsamples][n];  MODE 1 applies exp.
template <int MODE>
__global__ void k_transpose_out(int n, int ld, int nsamples, const double *__restrict__ src, double *__restrict__ dst)
{
    __shared__ double tile[32][33];
    const int row0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int row = row0 + j, s = s0 + threadIdx.x;
        tile[j][threadIdx.x] = (row < n && s < ld) ? src[(size_t)row * ld + s] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int s = s0 + j, row = row0 + threadIdx.x;
        if (s < nsamples && row < n) {
            double v = tile[threadIdx.x][j];
            if (MODE == 1) v = exp(v);
            dst[(size_t)s * n + row] = v;
        }
    }
}

// QoI: partial[(blockIdx.y)][s] = sum_rows obs[row] * x[row][s]   (DarcySolver::SolveFwd, src/DarcySolver.cpp:427)
__global__ void __launch_bounds__(TX *TY)
    k_dot_fixed(int n, int ld_, int rows_per_cta, const double *__restrict__ obs, const double *__restrict__ x,
                double *__restrict__ partial)
{
    __shared__ double2 red[TY][TX];
    const int sp = blockIdx.x * TX + threadIdx.x;
    const bool live = 2 * sp < ld_;
    const size_t so = 2 * (size_t)sp, ld = (size_t)ld_;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    double2 acc = make_double2(0.0, 0.0);
    if (live)
        for (int row = r0 + threadIdx.y; row < r1; row += TY) {
            const double w = __ldg(obs + row);
            if (w != 0.0) {
                const double2 xv = ld2(x + (size_t)row * ld + so);
                acc.x = fma(w, xv.x, acc.x);
                acc.y = fma(w, xv.y, acc.y);
            }
        }
    const double2 s = reduce_rows(acc, red);
    if (threadIdx.y == 0 && live) st2(partial + (size_t)blockIdx.y * ld + so, s);
}

__global__ void k_finish_sum(int ld, int nblk, const double *__restrict__ partial, double *__restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < ld) out[s] = sum_partials(partial, nblk, ld, s);
}

// Per-level moment sums of MLMC_Manager::InitRun (/root/reference/src/MLMC_Manager.cpp:123-131,:158-168),
// order {Y2, Y, ABSY, Q2, Q, ABSQ, C, Y3, Y4}.  One CTA, fixed reduction tree => deterministic.
// rows (nullable): [nsamples][4] = (Y, Q, Qc, C).  out9 is OVERWRITTEN with this batch's sums.
__global__ void __launch_bounds__(256)
    k_mlmc_accumulate(int nsamples, const double *__restrict__ Q, const double *__restrict__ Qc, double cost,
                      double *__restrict__ out9, double *__restrict__ rows)
{
    __shared__ double red[9][256];
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = threadIdx.x; j < nsamples; j += 256) {
        const double q = Q[j], qc = Qc ? Qc[j] : 0.0;
        const double y = Qc ? q - qc : q;
        a[7] += y * y * y;
        a[8] += y * y * y * y;
        a[0] += y * y;
        a[1] += y;
        a[2] += fabs(y);
        a[3] += q * q;
        a[4] += q;
        a[5] += fabs(q);
        a[6] += cost;
        if (rows) {
            rows[4 * (size_t)j + 0] = y;
            rows[4 * (size_t)j + 1] = q;
            rows[4 * (size_t)j + 2] = qc;
            rows[4 * (size_t)j + 3] = cost;
        }
    }
    for (int k = 0; k < 9; ++k) red[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int k = 0; k < 9; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 9) out9[threadIdx.x] = red[threadIdx.x][0];
}

__global__ void k_sum_int(int n, const int *__restrict__ v, unsigned long long *out)
{
    __shared__ unsigned long long red[256];
    unsigned long long a = 0;
    for (int j = threadIdx.x; j < n; j += 256) a += (unsigned long long)v[j];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out += red[0];
}

}  // namespace pmc
