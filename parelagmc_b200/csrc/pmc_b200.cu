// pmc_b200.cu -- context, host-once uploads, batched solver driver and the C ABI of include/pmc_b200.h.
//
// Product path: there is no CPU fallback in this file.  Every compute entry point launches the kernels of
// kernels.cuh / rng.cuh on the handle's stream and fails with PMC_ERR_CUDA if that is not possible.
#include "../../include/pmc_b200.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "host_sparse.hpp"
#include "kernels.cuh"
#include "rng.cuh"

namespace pmc {

// --------------------------------------------------------------------------------------------------
// device containers
// --------------------------------------------------------------------------------------------------
struct DevCsr {
    int rows = 0, cols = 0, nnz = 0;
    bool weighted = false;  // rowptr has 2*rows+1 entries, widx valid
    int *rowptr = nullptr, *col = nullptr, *widx = nullptr;
    double *val = nullptr;
    double matrix_bytes() const
    {
        return (double)nnz * (weighted ? 16.0 : 12.0) + (double)((weighted ? 2 : 1) * rows + 1) * 4.0;
    }
};

// host weighted CSR: per row a weighted segment then a fixed segment
struct HWCsr {
    int rows = 0, cols = 0;
    std::vector<int> rowptr2, col, widx;
    std::vector<double> val;
};
struct WEntry {
    int r, c, w;  // w < 0: fixed entry
    double v;
};
static HWCsr wcsr_from_entries(int rows, int cols, std::vector<WEntry> &e)
{
    // merge duplicates (same row, col, weight), weighted entries first within a row
    std::sort(e.begin(), e.end(), [](const WEntry &a, const WEntry &b) {
        if (a.r != b.r) return a.r < b.r;
        const int af = a.w < 0, bf = b.w < 0;
        if (af != bf) return af < bf;
        if (a.c != b.c) return a.c < b.c;
        return a.w < b.w;
    });
    HWCsr A;
    A.rows = rows;
    A.cols = cols;
    A.rowptr2.assign(2 * rows + 1, 0);
    std::vector<int> cntw(rows, 0), cntf(rows, 0);
    size_t i = 0;
    while (i < e.size()) {
        size_t j = i;
        double s = 0;
        while (j < e.size() && e[j].r == e[i].r && e[j].c == e[i].c && e[j].w == e[i].w) s += e[j++].v;
        if (s != 0.0) {
            A.col.push_back(e[i].c);
            A.widx.push_back(e[i].w < 0 ? 0 : e[i].w);
            A.val.push_back(s);
            (e[i].w < 0 ? cntf : cntw)[e[i].r]++;
        }
        i = j;
    }
    int p = 0;
    for (int r = 0; r < rows; ++r) {
        A.rowptr2[2 * r] = p;
        p += cntw[r];
        A.rowptr2[2 * r + 1] = p;
        p += cntf[r];
    }
    A.rowptr2[2 * rows] = p;
    return A;
}

struct VLevel {
    int n = 0;
    DevCsr S;                       // plain with values (sampler) or weighted with widx -> V_m (Darcy)
    DevCsr P, Pt;                   // to the next coarser V-level (absent on the last)
    double *l1inv_fixed = nullptr;  // sampler: [n]
    int nU = 0;                     // Darcy: number of distinct Schur values (diagonal + upper triangle)
    DevCsr T;                       // Darcy: V_m = T * V_{m-1}  (V_0 = T_0 * diag(M(k))^-1)
    DevCsr L;                       // Darcy: l1_m = L |V_m|
};

// Shape of the block-diagonal preconditioner of one system kind (sampler / Darcy).
struct PrecCfg {
    int mass_degree = 2;       // Chebyshev-Jacobi steps on the RT mass block
    int schur_degree = 2;      // Chebyshev steps of the V-cycle smoother on the Schur complement levels
    double schur_ratio = 4.0;  // smoother targets the eigenvalues in [1/ratio, 1] of the l1-scaled operator
    int coarse_degree = 8;     // Chebyshev steps on the coarsest V-cycle level
    double coarse_ratio = 30.0;
    double omega = 1.0;        // over-correction factor of the piecewise-constant coarse-grid correction
    int max_vlevels = 0;       // 0: as deep as the hierarchy allows; -1 (sampler): decide from the mass term
};

struct SaddleSys {
    bool ready = false, weighted = false;
    PrecCfg cfg;               // frozen at prepare time
    int *perm = nullptr;       // processing order of the block operator's rows (locality of the u-p coupling)
    int Nf = 0, Ne = 0, N = 0;
    DevCsr A;                       // block operator over N rows
    DevCsr Muu;                     // RT mass block (Nf rows)
    double *dinvM_fixed = nullptr;  // sampler: 1/diag(M)
    DevCsr Dm;                      // Darcy: diag M(k) = Dm * k_ext   (Nf x (Ne+1))
    double m_lo = 0.5, m_hi = 1.5;  // spectrum of diag(M)^-1 M
    std::vector<VLevel> v;
};

struct SamplerLevel {
    bool set = false, hasP = false;
    int Ne = 0, Nf = 0, lognormal = 1;
    double alpha = 0, g = 0;
    HCsr M, B, P;
    std::vector<double> Wdiag;
    double *w_sqrt = nullptr;
    DevCsr dP, dPt;
    SaddleSys sys;
};

struct DarcyLevel {
    bool set = false, hasP = false, ess_nonzero = false;
    int Ne = 0, Nf = 0;
    std::vector<int> elem_ptr, elem_dofs, ess_u;
    std::vector<double> elem_mat, ess_data, rhs, obs;
    HCsr B, Pp;
    double *d_rhs_bc = nullptr, *d_obs = nullptr, *d_ess_u_data = nullptr;
    DevCsr Mbc;  // weighted coupling of non-essential rows to essential columns (rhs fix-up)
    SaddleSys sys;
};

struct Arena {
    char *base = nullptr;
    size_t cap = 0, top = 0, peak = 0;
    bool dry = false, overflow = false;
    double *alloc(size_t count)
    {
        const size_t bytes = ((count * sizeof(double)) + 255) & ~(size_t)255;
        const size_t off = top;
        top += bytes;
        if (top > peak) peak = top;
        if (dry) return reinterpret_cast<double *>(off + 256);  // never dereferenced
        if (top > cap) { overflow = true; return nullptr; }
        return reinterpret_cast<double *>(base + off);
    }
    int *alloc_int(size_t count) { return reinterpret_cast<int *>(alloc((count + 1) / 2)); }
};

struct EventPair {
    cudaEvent_t a, b;
    int kclass;
};

}  // namespace pmc

using namespace pmc;

struct pmc_context_s {
    int device = 0, nlevels = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    double rel = 1e-6, abs_ = 1e-12;
    int maxit = 300;
    PrecCfg cfg_sampler, cfg_darcy;
    int max_batch = 0, check_every = 4;
    std::vector<SamplerLevel> s;
    std::vector<DarcyLevel> d;
    // rng
    bool rng_ready = false;
    double mu = 0.0, sigma = 1.0;
    RngTables *d_tab = nullptr;
    // memory
    Arena arena;
    std::vector<void *> owned;
    int *d_nactive = nullptr;
    int *h_nactive = nullptr;  // pinned
    unsigned long long *d_iters_total = nullptr;
    double *h_pinned = nullptr;  // pinned scratch for small results
    size_t h_pinned_count = 0;
    // stats
    unsigned profile_mask = 0;
    pmc_kernel_stats_t stats;
    std::vector<EventPair> ev_pending;
    std::vector<EventPair> ev_free;
    cudaError_t cuda_status = cudaSuccess;
};

typedef pmc_context_s Ctx;

static std::string g_create_error;

static int fail(Ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                               \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return fail(c, PMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// --------------------------------------------------------------------------------------------------
// launch helper with per-class statistics
// --------------------------------------------------------------------------------------------------
static void ev_begin(Ctx *c, int kclass, EventPair &ep, bool &timed)
{
    timed = (c->profile_mask >> kclass) & 1u;
    if (!timed) return;
    if (!c->ev_free.empty()) {
        ep = c->ev_free.back();
        c->ev_free.pop_back();
    } else {
        cudaEventCreate(&ep.a);
        cudaEventCreate(&ep.b);
    }
    ep.kclass = kclass;
    cudaEventRecord(ep.a, c->stream);
}
static void ev_end(Ctx *c, EventPair &ep, bool timed)
{
    if (!timed) return;
    cudaEventRecord(ep.b, c->stream);
    c->ev_pending.push_back(ep);
}

template <typename... KArgs, typename... Args>
static void launch(Ctx *c, int kclass, double bytes, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args)
{
    EventPair ep;
    bool timed;
    ev_begin(c, kclass, ep, timed);
    kernel<<<grid, block, 0, c->stream>>>(args...);
    ev_end(c, ep, timed);
    c->stats.launches[kclass]++;
    c->stats.algo_bytes[kclass] += bytes;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess && c->cuda_status == cudaSuccess) c->cuda_status = e;
}

static void resolve_events(Ctx *c)
{
    if (c->ev_pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto &ep : c->ev_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) {
            c->stats.ms[ep.kclass] += ms;
            c->stats.timed_launches[ep.kclass]++;
        }
        c->ev_free.push_back(ep);
    }
    c->ev_pending.clear();
}

// --------------------------------------------------------------------------------------------------
// grid shapes
// --------------------------------------------------------------------------------------------------
struct Shape {
    dim3 grid, block;
    int rows_per_cta, nblk;
};
static Shape shape_for(int n, int ld)
{
    Shape s;
    const int sblocks = (ld / 2 + TX - 1) / TX;
    int want = (148 * 16 + sblocks - 1) / sblocks;  // ~2 waves of 8 resident CTAs per SM
    int maxblk = (n + TY - 1) / TY;
    if (maxblk < 1) maxblk = 1;
    if (want > 512) want = 512;
    if (want > maxblk) want = maxblk;
    if (want < 1) want = 1;
    int rpc = (n + want - 1) / want;
    rpc = ((rpc + TY - 1) / TY) * TY;
    if (rpc < TY) rpc = TY;
    s.rows_per_cta = rpc;
    s.nblk = n > 0 ? (n + rpc - 1) / rpc : 1;
    s.grid = dim3(sblocks, s.nblk);
    s.block = dim3(TX, TY);
    return s;
}

// --------------------------------------------------------------------------------------------------
// uploads
// --------------------------------------------------------------------------------------------------
template <typename T>
static int to_device(Ctx *c, const std::vector<T> &h, T **out)
{
    *out = nullptr;
    const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes));
    c->owned.push_back(p);
    if (!h.empty()) CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<T *>(p);
    return PMC_OK;
}

static int upload_csr(Ctx *c, const HCsr &A, DevCsr &D)
{
    D.rows = A.rows;
    D.cols = A.cols;
    D.nnz = A.nnz();
    D.weighted = false;
    int rc;
    if ((rc = to_device(c, A.rowptr, &D.rowptr))) return rc;
    if ((rc = to_device(c, A.col, &D.col))) return rc;
    if ((rc = to_device(c, A.val, &D.val))) return rc;
    return PMC_OK;
}

static int upload_wcsr(Ctx *c, const HWCsr &A, DevCsr &D)
{
    D.rows = A.rows;
    D.cols = A.cols;
    D.nnz = (int)A.col.size();
    D.weighted = true;
    int rc;
    if ((rc = to_device(c, A.rowptr2, &D.rowptr))) return rc;
    if ((rc = to_device(c, A.col, &D.col))) return rc;
    if ((rc = to_device(c, A.widx, &D.widx))) return rc;
    if ((rc = to_device(c, A.val, &D.val))) return rc;
    return PMC_OK;
}

// Unique-value numbering of a symmetric pattern: uid(i,j) = uid(j,i), diagonal included.
struct SymPattern {
    int n = 0, nU = 0;
    std::vector<int> rowptr, col, uid;  // full pattern (both triangles)
    int find(int i, int j) const
    {
        const int *b = col.data() + rowptr[i], *e = col.data() + rowptr[i + 1];
        const int *p = std::lower_bound(b, e, j);
        if (p == e || *p != j) return -1;
        return uid[p - col.data()];
    }
};

static SymPattern sym_pattern_from(const HCsr &S)
{
    SymPattern sp;
    sp.n = S.rows;
    sp.rowptr = S.rowptr;
    sp.col = S.col;
    sp.uid.assign(S.col.size(), -1);
    int next = 0;
    for (int i = 0; i < S.rows; ++i)
        for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p)
            if (S.col[p] >= i) sp.uid[p] = next++;
    sp.nU = next;
    for (int i = 0; i < S.rows; ++i)
        for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p)
            if (S.col[p] < i) {
                const int j = S.col[p];
                const int *b = S.col.data() + S.rowptr[j], *e = S.col.data() + S.rowptr[j + 1];
                const int *q = std::lower_bound(b, e, i);
                sp.uid[p] = (q != e && *q == i) ? sp.uid[q - S.col.data()] : -1;
            }
    return sp;
}

static HCsr symmetrize_pattern(const HCsr &S)
{
    std::vector<Coo> e;
    e.reserve(2 * S.col.size());
    for (int i = 0; i < S.rows; ++i)
        for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p) {
            e.push_back({i, S.col[p], 1.0});
            e.push_back({S.col[p], i, 1.0});
        }
    for (int i = 0; i < S.rows; ++i) e.push_back({i, i, 1.0});
    return csr_from_coo(S.rows, S.rows, e);
}

// Processing order of the saddle operator's rows: every element row (Nf + e) is preceded by the face rows it
// "owns" (faces whose lowest-numbered adjacent element is e).  Storage order is unchanged; a CTA's row block then
// touches u and p entries that were fetched recently, so the gathers hit L2 instead of re-reading HBM.
static std::vector<int> saddle_row_order(const HCsr &B, int Nf, int Ne)
{
    std::vector<int> owner(Nf, -1);
    for (int e = 0; e < Ne; ++e)
        for (int p = B.rowptr[e]; p < B.rowptr[e + 1]; ++p) {
            const int f = B.col[p];
            if (f >= 0 && f < Nf && owner[f] < 0) owner[f] = e;
        }
    std::vector<int> cnt(Ne + 2, 0);
    for (int f = 0; f < Nf; ++f) cnt[(owner[f] < 0 ? Ne : owner[f]) + 1]++;
    for (int e = 0; e <= Ne; ++e) cnt[e + 1] += cnt[e];
    std::vector<int> faces(Nf), pos(cnt.begin(), cnt.end() - 1);
    for (int f = 0; f < Nf; ++f) faces[pos[owner[f] < 0 ? Ne : owner[f]]++] = f;
    std::vector<int> perm;
    perm.reserve(Nf + Ne);
    for (int e = 0; e < Ne; ++e) {
        for (int q = cnt[e]; q < cnt[e + 1]; ++q) perm.push_back(faces[q]);
        perm.push_back(Nf + e);
    }
    for (int q = cnt[Ne]; q < cnt[Ne + 1]; ++q) perm.push_back(faces[q]);
    return perm;
}

// ---- sampler system (everything fixed across samples) ---------------------------------------------------
static int prepare_sampler(Ctx *c, int level)
{
    SamplerLevel &L = c->s[level];
    SaddleSys &sys = L.sys;
    if (sys.ready) return PMC_OK;
    const int Nf = L.Nf, Ne = L.Ne, N = Nf + Ne;
    sys.weighted = false;
    sys.Nf = Nf;
    sys.Ne = Ne;
    sys.N = N;
    sys.cfg = c->cfg_sampler;
    {
        std::vector<int> perm = saddle_row_order(L.B, Nf, Ne);
        int rcp = to_device(c, perm, &sys.perm);
        if (rcp) return rcp;
    }
    HCsr Bt = csr_transpose(L.B);
    {  // block operator [[M, B^T], [B, -alpha W]]   (/root/reference/src/PDESampler.cpp:279-284)
        std::vector<Coo> e;
        e.reserve(L.M.nnz() + 2 * L.B.nnz() + Ne);
        for (int i = 0; i < Nf; ++i) {
            for (int p = L.M.rowptr[i]; p < L.M.rowptr[i + 1]; ++p) e.push_back({i, L.M.col[p], L.M.val[p]});
            for (int p = Bt.rowptr[i]; p < Bt.rowptr[i + 1]; ++p) e.push_back({i, Nf + Bt.col[p], Bt.val[p]});
        }
        for (int i = 0; i < Ne; ++i) {
            for (int p = L.B.rowptr[i]; p < L.B.rowptr[i + 1]; ++p) e.push_back({Nf + i, L.B.col[p], L.B.val[p]});
            e.push_back({Nf + i, Nf + i, -1.0 * L.alpha * L.Wdiag[i]});  // W_s <- -alpha W_s (:256-258)
        }
        HCsr A = csr_from_coo(N, N, e);
        int rc = upload_csr(c, A, sys.A);
        if (rc) return rc;
    }
    int rc = upload_csr(c, L.M, sys.Muu);
    if (rc) return rc;
    std::vector<double> Md = csr_diag(L.M);
    std::vector<double> dinv(Nf);
    for (int i = 0; i < Nf; ++i) dinv[i] = Md[i] != 0.0 ? 1.0 / Md[i] : 1.0;
    if ((rc = to_device(c, dinv, &sys.dinvM_fixed))) return rc;
    sys.m_hi = 1.02 * lambda_max_scaled(L.M, Md);
    sys.m_lo = sys.m_hi / 3.0;
    // Schur complement S = alpha W + B diag(M)^-1 B^T and its Galerkin hierarchy on the sampler's own P
    HCsr Bs = L.B;
    for (int i = 0; i < Ne; ++i)
        for (int p = Bs.rowptr[i]; p < Bs.rowptr[i + 1]; ++p) Bs.val[p] *= dinv[Bs.col[p]];
    HCsr S = csr_matmul(Bs, Bt);
    {
        std::vector<Coo> e;
        for (int i = 0; i < Ne; ++i) {
            for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p) e.push_back({i, S.col[p], S.val[p]});
            e.push_back({i, i, L.alpha * L.Wdiag[i]});
        }
        S = csr_from_coo(Ne, Ne, e);
    }
    std::vector<const HCsr *> Ps;
    for (int m = level; m < c->nlevels - 1 && c->s[m].set && c->s[m].hasP; ++m) Ps.push_back(&c->s[m].P);
    if (sys.cfg.max_vlevels < 0) {
        // lambda_min of the l1-scaled Schur complement is at least min_e alpha W_e / l1_e: when that is not small the
        // -alpha W block makes S well conditioned and one Chebyshev sweep over [lambda_min, 1] replaces the V-cycle
        double lmin = 1.0;
        for (int i = 0; i < Ne; ++i) {
            double t = 0;
            for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p) t += std::fabs(S.val[p]);
            if (t > 0) lmin = std::min(lmin, L.alpha * L.Wdiag[i] / t);
        }
        if (lmin >= 1.0 / 20.0) {
            sys.cfg.max_vlevels = 1;
            sys.cfg.coarse_ratio = std::max(2.0, 1.0 / lmin);
            sys.cfg.coarse_degree = sys.cfg.coarse_ratio <= 8.0 ? 3 : 4;
        } else
            sys.cfg.max_vlevels = 0;
    }
    if (sys.cfg.max_vlevels > 0 && (int)Ps.size() > sys.cfg.max_vlevels - 1) Ps.resize(sys.cfg.max_vlevels - 1);
    sys.v.resize(Ps.size() + 1);
    for (size_t m = 0; m < sys.v.size(); ++m) {
        VLevel &V = sys.v[m];
        V.n = S.rows;
        if ((rc = upload_csr(c, S, V.S))) return rc;
        std::vector<double> l1(S.rows);
        for (int i = 0; i < S.rows; ++i) {
            double t = 0;
            for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p) t += std::fabs(S.val[p]);
            l1[i] = t > 0 ? 1.0 / t : 1.0;
        }
        if ((rc = to_device(c, l1, &V.l1inv_fixed))) return rc;
        if (m < Ps.size()) {
            const HCsr &P = *Ps[m];
            HCsr Pt = csr_transpose(P);
            if ((rc = upload_csr(c, P, V.P))) return rc;
            if ((rc = upload_csr(c, Pt, V.Pt))) return rc;
            S = csr_matmul(Pt, csr_matmul(S, P));
        }
    }
    sys.ready = true;
    return PMC_OK;
}

// ---- Darcy system (values depend on the sample through k) -----------------------------------------------
static int prepare_darcy(Ctx *c, int level)
{
    DarcyLevel &L = c->d[level];
    SaddleSys &sys = L.sys;
    if (sys.ready) return PMC_OK;
    const int Nf = L.Nf, Ne = L.Ne, N = Nf + Ne;
    sys.weighted = true;
    sys.Nf = Nf;
    sys.Ne = Ne;
    sys.N = N;
    sys.cfg = c->cfg_darcy;
    {
        std::vector<int> perm = saddle_row_order(L.B, Nf, Ne);
        int rcp = to_device(c, perm, &sys.perm);
        if (rcp) return rcp;
    }
    const std::vector<int> &ess = L.ess_u;
    // Be: essential columns removed
    HCsr Be;
    {
        std::vector<Coo> e;
        for (int i = 0; i < Ne; ++i)
            for (int p = L.B.rowptr[i]; p < L.B.rowptr[i + 1]; ++p)
                if (!ess[L.B.col[p]]) e.push_back({i, L.B.col[p], L.B.val[p]});
        Be = csr_from_coo(Ne, Nf, e, true);
    }
    HCsr Bet = csr_transpose(Be);
    // element triples of M(k) = sum_e k_e R_e^T M_e R_e  (/root/reference/src/DarcySolver.cpp:479), with
    // EliminateRowCol on the essential dofs (:495-498) folded in: essential rows become identity rows,
    // essential columns are dropped from the operator and kept in Mbc for the right-hand-side fix-up.
    std::vector<WEntry> eM, eA, eBc;
    std::vector<Coo> eDm;
    double lam_hi = 1.0;
    {
        size_t off = 0;
        for (int el = 0; el < Ne; ++el) {
            const int n = L.elem_ptr[el + 1] - L.elem_ptr[el];
            const int *dof = L.elem_dofs.data() + L.elem_ptr[el];
            const double *Me = L.elem_mat.data() + off;
            lam_hi = std::max(lam_hi, dense_lambda_max_scaled(Me, n));
            for (int a = 0; a < n; ++a) {
                const int i = dof[a];
                if (!ess[i]) eDm.push_back({i, el, Me[a * n + a]});
                for (int b = 0; b < n; ++b) {
                    const int j = dof[b];
                    const double v = Me[a * n + b];
                    if (v == 0.0 || ess[i]) continue;
                    if (ess[j]) eBc.push_back({i, j, el, v});
                    else {
                        eM.push_back({i, j, el, v});
                        eA.push_back({i, j, el, v});
                    }
                }
            }
            off += (size_t)n * n;
        }
    }
    for (int i = 0; i < Nf; ++i)
        if (ess[i]) {
            eM.push_back({i, i, -1, 1.0});
            eA.push_back({i, i, -1, 1.0});
            eDm.push_back({i, Ne, 1.0});  // column Ne of k_ext is the constant 1
        }
    for (int i = 0; i < Nf; ++i)
        for (int p = Bet.rowptr[i]; p < Bet.rowptr[i + 1]; ++p) eA.push_back({i, Nf + Bet.col[p], -1, Bet.val[p]});
    for (int i = 0; i < Ne; ++i)
        for (int p = Be.rowptr[i]; p < Be.rowptr[i + 1]; ++p) eA.push_back({Nf + i, Be.col[p], -1, Be.val[p]});
    int rc;
    {
        HWCsr A = wcsr_from_entries(N, N, eA);
        if ((rc = upload_wcsr(c, A, sys.A))) return rc;
        HWCsr M = wcsr_from_entries(Nf, Nf, eM);
        if ((rc = upload_wcsr(c, M, sys.Muu))) return rc;
        HWCsr Mbc = wcsr_from_entries(Nf, Nf, eBc);
        if ((rc = upload_wcsr(c, Mbc, L.Mbc))) return rc;
        HCsr Dm = csr_from_coo(Nf, Ne + 1, eDm);
        if ((rc = upload_csr(c, Dm, sys.Dm))) return rc;
    }
    sys.m_hi = lam_hi;
    sys.m_lo = lam_hi / 3.0;
    // right-hand side after elimination (sample-independent part)
    {
        std::vector<double> b(L.rhs);
        std::vector<double> eu(Nf, 0.0);
        L.ess_nonzero = false;
        for (int j = 0; j < Nf; ++j)
            if (ess[j]) {
                eu[j] = L.ess_data[j];
                if (eu[j] != 0.0) L.ess_nonzero = true;
            }
        for (int i = 0; i < Ne; ++i)
            for (int p = L.B.rowptr[i]; p < L.B.rowptr[i + 1]; ++p)
                if (ess[L.B.col[p]]) b[Nf + i] -= L.B.val[p] * L.ess_data[L.B.col[p]];
        for (int j = 0; j < Nf; ++j)
            if (ess[j]) b[j] = L.ess_data[j];
        if ((rc = to_device(c, b, &L.d_rhs_bc))) return rc;
        if ((rc = to_device(c, eu, &L.d_ess_u_data))) return rc;
        if ((rc = to_device(c, L.obs, &L.d_obs))) return rc;
    }
    // Schur complement S(k) = Be diag(M(k))^-1 Be^T: pattern, unique-value map T_0 and the Galerkin chain
    std::vector<const HCsr *> Ps;
    for (int m = level; m < c->nlevels - 1 && c->d[m].set && c->d[m].hasP; ++m) Ps.push_back(&c->d[m].Pp);
    if (sys.cfg.max_vlevels > 0 && (int)Ps.size() > sys.cfg.max_vlevels - 1) Ps.resize(sys.cfg.max_vlevels - 1);
    sys.v.resize(Ps.size() + 1);
    HCsr Spat;
    {
        HCsr Bo = Be, Bto = Bet;
        for (auto &v : Bo.val) v = 1.0;
        for (auto &v : Bto.val) v = 1.0;
        Spat = symmetrize_pattern(csr_matmul(Bo, Bto));
    }
    SymPattern sp = sym_pattern_from(Spat);
    {  // T_0[uid(e,e')][f] = b_ef b_e'f  for e <= e'
        std::vector<Coo> e;
        for (int f = 0; f < Nf; ++f)
            for (int p = Bet.rowptr[f]; p < Bet.rowptr[f + 1]; ++p)
                for (int q = Bet.rowptr[f]; q < Bet.rowptr[f + 1]; ++q) {
                    const int i = Bet.col[p], j = Bet.col[q];
                    if (i <= j) e.push_back({sp.find(i, j), f, Bet.val[p] * Bet.val[q]});
                }
        HCsr T0 = csr_from_coo(sp.nU, Nf, e);
        if ((rc = upload_csr(c, T0, sys.v[0].T))) return rc;
    }
    for (size_t m = 0; m < sys.v.size(); ++m) {
        VLevel &V = sys.v[m];
        V.n = sp.n;
        V.nU = sp.nU;
        {  // S_m as weighted CSR over V_m (coefficient 1) and the l1 map
            std::vector<WEntry> e;
            std::vector<Coo> el;
            for (int i = 0; i < sp.n; ++i)
                for (int p = sp.rowptr[i]; p < sp.rowptr[i + 1]; ++p) {
                    e.push_back({i, sp.col[p], sp.uid[p], 1.0});
                    el.push_back({i, sp.uid[p], 1.0});
                }
            HWCsr Sw = wcsr_from_entries(sp.n, sp.n, e);
            if ((rc = upload_wcsr(c, Sw, V.S))) return rc;
            HCsr Lm = csr_from_coo(sp.n, sp.nU, el);
            if ((rc = upload_csr(c, Lm, V.L))) return rc;
        }
        if (m < Ps.size()) {
            const HCsr &P = *Ps[m];
            HCsr Pt = csr_transpose(P);
            if ((rc = upload_csr(c, P, V.P))) return rc;
            if ((rc = upload_csr(c, Pt, V.Pt))) return rc;
            // coarse pattern and T_{m+1}: S_c[I,J] = sum_{ij} P_iI S_ij P_jJ
            HCsr Sone;
            Sone.rows = Sone.cols = sp.n;
            Sone.rowptr = sp.rowptr;
            Sone.col = sp.col;
            Sone.val.assign(sp.col.size(), 1.0);
            HCsr Pabs = P;
            for (auto &v : Pabs.val) v = 1.0;
            HCsr Pabst = csr_transpose(Pabs);
            HCsr Cpat = symmetrize_pattern(csr_matmul(Pabst, csr_matmul(Sone, Pabs)));
            SymPattern spc = sym_pattern_from(Cpat);
            std::vector<Coo> e;
            for (int i = 0; i < sp.n; ++i)
                for (int p = sp.rowptr[i]; p < sp.rowptr[i + 1]; ++p) {
                    const int j = sp.col[p];
                    for (int a = P.rowptr[i]; a < P.rowptr[i + 1]; ++a)
                        for (int b = P.rowptr[j]; b < P.rowptr[j + 1]; ++b) {
                            const int I = P.col[a], J = P.col[b];
                            if (I <= J) e.push_back({spc.find(I, J), sp.uid[p], P.val[a] * P.val[b]});
                        }
                }
            HCsr T = csr_from_coo(spc.nU, sp.nU, e);
            if ((rc = upload_csr(c, T, sys.v[m + 1].T))) return rc;
            sp = std::move(spc);
        }
    }
    sys.ready = true;
    return PMC_OK;
}

// --------------------------------------------------------------------------------------------------
// solve workspace
// --------------------------------------------------------------------------------------------------
struct SolveWs {
    int ld = 0;
    double *v0, *v1, *w0, *w1, *u1, *q, *x, *b;  // MINRES, N x ld each
    double *mu_d, *mu_z;                         // mass-block Chebyshev scratch, Nf x ld
    double *dinvM;                               // Darcy: batched 1/diag M(k), Nf x ld
    std::vector<double *> vr, vzA, vzB, vd, vres, vV, vl1;  // per V-level
    double *partial;                                          // [npart][ld]
    double *st;                                               // ST_COUNT x ld
    int *active, *iters;
    int npart = 0;
};

static void carve_solve(Arena &ar, const SaddleSys &sys, int ld, SolveWs &ws)
{
    const size_t N = sys.N, Nf = sys.Nf, S = ld;
    ws.ld = ld;
    ws.v0 = ar.alloc(N * S); ws.v1 = ar.alloc(N * S); ws.w0 = ar.alloc(N * S); ws.w1 = ar.alloc(N * S);
    ws.u1 = ar.alloc(N * S); ws.q = ar.alloc(N * S); ws.x = ar.alloc(N * S); ws.b = ar.alloc(N * S);
    ws.mu_d = ar.alloc(Nf * S);
    ws.mu_z = ar.alloc(Nf * S);
    ws.dinvM = sys.weighted ? ar.alloc(Nf * S) : nullptr;
    const size_t nv = sys.v.size();
    ws.vr.assign(nv, nullptr); ws.vzA.assign(nv, nullptr); ws.vzB.assign(nv, nullptr); ws.vd.assign(nv, nullptr);
    ws.vres.assign(nv, nullptr); ws.vV.assign(nv, nullptr); ws.vl1.assign(nv, nullptr);
    for (size_t m = 0; m < nv; ++m) {
        const size_t n = sys.v[m].n;
        if (m > 0) { ws.vr[m] = ar.alloc(n * S); ws.vzA[m] = ar.alloc(n * S); }
        ws.vzB[m] = ar.alloc(n * S);
        ws.vd[m] = ar.alloc(n * S);
        if (m + 1 < nv) ws.vres[m] = ar.alloc(n * S);
        if (sys.weighted) { ws.vV[m] = ar.alloc((size_t)sys.v[m].nU * S); ws.vl1[m] = ar.alloc(n * S); }
    }
    ws.npart = shape_for(sys.N, ld).nblk + shape_for(sys.Nf, ld).nblk + shape_for(sys.Ne, ld).nblk + 4;
    ws.partial = ar.alloc((size_t)ws.npart * S);
    ws.st = ar.alloc((size_t)ST_COUNT * S);
    ws.active = ar.alloc_int(S);
    ws.iters = ar.alloc_int(S);
}

static int ensure_arena(Ctx *c, size_t bytes)
{
    if (c->arena.cap >= bytes) return PMC_OK;
    if (c->arena.base) {
        cudaStreamSynchronize(c->stream);
        cudaFree(c->arena.base);
        c->arena.base = nullptr;
        c->arena.cap = 0;
    }
    const size_t want = bytes + (bytes >> 4) + (1 << 20);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(c, PMC_ERR_NOMEM, "cudaMalloc of %zu bytes of batch workspace failed: %s", want, cudaGetErrorString(e));
    }
    c->arena.base = (char *)p;
    c->arena.cap = want;
    return PMC_OK;
}

// --------------------------------------------------------------------------------------------------
// operator applies
// --------------------------------------------------------------------------------------------------
static double spmm_bytes(const DevCsr &A, int ld, double vec_rows)
{
    return A.matrix_bytes() + (double)ld * vec_rows * 8.0;
}

// generic fused SpMM dispatch
static void spmm(Ctx *c, int kclass, int ep, const DevCsr &A, const double *V, int ld, const double *x, double *y,
                 const double *r, double *d, const double *dinv, bool bdinv, double ca, double cb, const double *dotw,
                 double *partial, int partial_off, double vec_rows, const int *perm = nullptr)
{
    const Shape sh = shape_for(A.rows, ld);
    SpmmArgs a;
    a.perm = perm;
    a.n = A.rows; a.ld = ld; a.rows_per_cta = sh.rows_per_cta;
    a.rowptr = A.rowptr; a.col = A.col; a.val = A.val; a.widx = A.widx; a.V = V;
    a.x = x; a.y = y; a.r = r; a.d = d; a.dinv = dinv; a.ca = ca; a.cb = cb;
    a.dotw = dotw; a.partial = partial; a.partial_off = partial_off;
    const bool dot = dotw != nullptr;
    const double bytes = spmm_bytes(A, ld, vec_rows);
    const bool w = A.weighted;
#define PMC_SPMM_CASE(EP_, W_, BD_, DOT_) \
    launch(c, kclass, bytes, k_spmm<EP_, W_, BD_, DOT_>, sh.grid, sh.block, a)
    if (ep == EP_AX) {
        if (w) { if (dot) PMC_SPMM_CASE(EP_AX, true, false, true); else PMC_SPMM_CASE(EP_AX, true, false, false); }
        else   { if (dot) PMC_SPMM_CASE(EP_AX, false, false, true); else PMC_SPMM_CASE(EP_AX, false, false, false); }
    } else if (ep == EP_RESID) {
        if (w) PMC_SPMM_CASE(EP_RESID, true, false, false); else PMC_SPMM_CASE(EP_RESID, false, false, false);
    } else if (ep == EP_ADD) {
        if (w) PMC_SPMM_CASE(EP_ADD, true, false, false); else PMC_SPMM_CASE(EP_ADD, false, false, false);
    } else {
        if (w) { if (dot) PMC_SPMM_CASE(EP_CHEB, true, true, true); else PMC_SPMM_CASE(EP_CHEB, true, true, false); }
        else   { if (dot) PMC_SPMM_CASE(EP_CHEB, false, false, true); else PMC_SPMM_CASE(EP_CHEB, false, false, false); }
        (void)bdinv;
    }
#undef PMC_SPMM_CASE
}

struct ChebOp {
    const DevCsr *A;
    const double *V;     // weights (weighted operators)
    const double *dinv;  // fixed [n] or batched [n][ld] (batched iff A->weighted)
    double lo, hi;
    double vrows;        // weight rows read per apply (for the byte count)
    int kclass;
};

// `deg` Chebyshev steps for A z = r.  from_zero: z_0 = 0, the result ends in `end_buf`.  Otherwise the current
// iterate lives in `cur` and the result ends in (deg even ? cur : other).  Returns the buffer holding the result.
static double *cheb_run(Ctx *c, const ChebOp &op, int ld, const double *r, double *d, int deg, bool from_zero,
                        double *cur, double *other, bool dot, double *partial, int partial_off)
{
    const int n = op.A->rows;
    const bool bd = op.A->weighted;
    const double theta = 0.5 * (op.hi + op.lo), delta = 0.5 * (op.hi - op.lo), sigma = theta / delta;
    double rho = 1.0 / sigma;
    double *zin = cur, *zout = other;
    int j0 = 0;
    if (from_zero) {
        // writes alternate; the last of `deg` writes must hit end_buf == cur
        zout = (deg % 2 == 1) ? cur : other;
        const Shape sh = shape_for(n, ld);
        const bool d0 = dot && deg == 1;
        const double bytes = (double)ld * n * (3 + (bd ? 1 : 0)) * 8.0;
        if (bd) {
            if (d0) launch(c, op.kclass, bytes, k_cheb_first<true, true>, sh.grid, sh.block, n, ld, sh.rows_per_cta, r, op.dinv, 1.0 / theta, d, zout, partial, partial_off);
            else launch(c, op.kclass, bytes, k_cheb_first<true, false>, sh.grid, sh.block, n, ld, sh.rows_per_cta, r, op.dinv, 1.0 / theta, d, zout, partial, partial_off);
        } else {
            if (d0) launch(c, op.kclass, bytes, k_cheb_first<false, true>, sh.grid, sh.block, n, ld, sh.rows_per_cta, r, op.dinv, 1.0 / theta, d, zout, partial, partial_off);
            else launch(c, op.kclass, bytes, k_cheb_first<false, false>, sh.grid, sh.block, n, ld, sh.rows_per_cta, r, op.dinv, 1.0 / theta, d, zout, partial, partial_off);
        }
        zin = zout;
        zout = (zin == cur) ? other : cur;
        j0 = 1;
    }
    for (int j = j0; j < deg; ++j) {
        double ca, cb;
        if (j == 0) { ca = 0.0; cb = 1.0 / theta; }
        else {
            const double rho_new = 1.0 / (2.0 * sigma - rho);
            ca = rho_new * rho;
            cb = 2.0 * rho_new / delta;
            rho = rho_new;
        }
        const bool dl = dot && j == deg - 1;
        spmm(c, op.kclass, EP_CHEB, *op.A, op.V, ld, zin, zout, r, d, op.dinv, bd, ca, cb, dl ? r : nullptr, partial,
             partial_off, n * (5.0 + (bd ? 1 : 0)) + op.vrows);
        double *t = zin; zin = zout; zout = t;
    }
    return zin;
}

struct Solver {
    Ctx *c;
    SaddleSys *sys;
    SolveWs *ws;
    const double *k_ext;  // Darcy weights [Ne+1][ld] or null
    int ld;
};

static void vcycle(Solver &sv, int m, const double *r, double *zout, double *ztmp, bool dot, int partial_off)
{
    Ctx *c = sv.c;
    SaddleSys &sys = *sv.sys;
    SolveWs &ws = *sv.ws;
    VLevel &L = sys.v[m];
    const int ld = sv.ld;
    ChebOp op;
    op.A = &L.S;
    op.V = sys.weighted ? ws.vV[m] : nullptr;
    op.dinv = sys.weighted ? ws.vl1[m] : L.l1inv_fixed;
    op.hi = 1.0;
    op.vrows = sys.weighted ? L.nU : 0;
    op.kclass = PMC_K_SCHUR_SMOOTH;
    const bool last = (m + 1 == (int)sys.v.size());
    const PrecCfg &cfg = sys.cfg;
    if (last) {
        op.lo = 1.0 / cfg.coarse_ratio;
        cheb_run(c, op, ld, r, ws.vd[m], cfg.coarse_degree, true, zout, ztmp, dot, ws.partial, partial_off);
        return;
    }
    op.lo = 1.0 / cfg.schur_ratio;
    const int s = cfg.schur_degree;
    double *E = (s % 2 == 0) ? zout : ztmp, *O = (s % 2 == 0) ? ztmp : zout;
    cheb_run(c, op, ld, r, ws.vd[m], s, true, E, O, false, nullptr, 0);
    spmm(c, PMC_K_SCHUR_SMOOTH, EP_RESID, L.S, op.V, ld, E, ws.vres[m], r, nullptr, nullptr, false, 0, 0, nullptr,
         nullptr, 0, 3.0 * L.n + op.vrows);
    spmm(c, PMC_K_TRANSFER, EP_AX, L.Pt, nullptr, ld, ws.vres[m], ws.vr[m + 1], nullptr, nullptr, nullptr, false, 0, 0,
         nullptr, nullptr, 0, (double)L.n + sys.v[m + 1].n);
    vcycle(sv, m + 1, ws.vr[m + 1], ws.vzA[m + 1], ws.vzB[m + 1], false, 0);
    spmm(c, PMC_K_TRANSFER, EP_ADD, L.P, nullptr, ld, ws.vzA[m + 1], E, nullptr, nullptr, nullptr, false, cfg.omega, 0,
         nullptr, nullptr, 0, 2.0 * L.n + sys.v[m + 1].n);
    cheb_run(c, op, ld, r, ws.vd[m], s, false, E, O, dot, ws.partial, partial_off);
}

// z = Prec r (block diagonal); if dot, partial[0..nblk) receives the row-block partial sums of r.z.  Returns nblk.
static int apply_prec(Solver &sv, const double *r, double *z, bool dot)
{
    Ctx *c = sv.c;
    SaddleSys &sys = *sv.sys;
    SolveWs &ws = *sv.ws;
    const int ld = sv.ld;
    const size_t po = (size_t)sys.Nf * ld;
    ChebOp op;
    op.A = &sys.Muu;
    op.V = sv.k_ext;
    op.dinv = sys.weighted ? ws.dinvM : sys.dinvM_fixed;
    op.lo = sys.m_lo;
    op.hi = sys.m_hi;
    op.vrows = sys.weighted ? sys.Ne : 0;
    op.kclass = PMC_K_MASS_SMOOTH;
    cheb_run(c, op, ld, r, ws.mu_d, sys.cfg.mass_degree, true, z, ws.mu_z, dot, ws.partial, 0);
    const int nbu = shape_for(sys.Nf, ld).nblk;
    vcycle(sv, 0, r + po, z + po, ws.vzB[0], dot, nbu);
    return nbu + shape_for(sys.Ne, ld).nblk;
}

static void saddle_apply(Solver &sv, int ep, const double *x, double *y, const double *r, bool dot, int kclass)
{
    SaddleSys &sys = *sv.sys;
    spmm(sv.c, kclass, ep, sys.A, sv.k_ext, sv.ld, x, y, r, nullptr, nullptr, false, 0, 0, dot ? x : nullptr,
         sv.ws->partial, 0, (ep == EP_RESID ? 3.0 : 2.0) * sys.N + (sys.weighted ? sys.Ne : 0), sys.perm);
}

static void fill(Ctx *c, double *p, size_t n, double v)
{
    if (n == 0) return;
    int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    launch(c, PMC_K_MISC, (double)n * 8.0, k_fill, dim3(blocks), dim3(256), p, n, v);
}

// Darcy per-solve values: 1/diag M(k), Schur values on every V-level, l1 norms.
static void darcy_setup_values(Solver &sv)
{
    Ctx *c = sv.c;
    SaddleSys &sys = *sv.sys;
    SolveWs &ws = *sv.ws;
    const int ld = sv.ld;
    {
        const Shape sh = shape_for(sys.Nf, ld);
        launch(c, PMC_K_SETUP, spmm_bytes(sys.Dm, ld, sys.Ne + sys.Nf), k_spmm_setup<false, true>, sh.grid, sh.block,
               sys.Nf, ld, sh.rows_per_cta, sys.Dm.rowptr, sys.Dm.col, sys.Dm.val, sv.k_ext, ws.dinvM);
    }
    const double *prev = ws.dinvM;
    int prev_rows = sys.Nf;
    for (size_t m = 0; m < sys.v.size(); ++m) {
        VLevel &L = sys.v[m];
        Shape sh = shape_for(L.nU, ld);
        launch(c, PMC_K_SETUP, spmm_bytes(L.T, ld, prev_rows + L.nU), k_spmm_setup<false, false>, sh.grid, sh.block,
               L.nU, ld, sh.rows_per_cta, L.T.rowptr, L.T.col, L.T.val, prev, ws.vV[m]);
        sh = shape_for(L.n, ld);
        launch(c, PMC_K_SETUP, spmm_bytes(L.L, ld, L.nU + L.n), k_spmm_setup<true, true>, sh.grid, sh.block, L.n, ld,
               sh.rows_per_cta, L.L.rowptr, L.L.col, L.L.val, ws.vV[m], ws.vl1[m]);
        prev = ws.vV[m];
        prev_rows = L.nU;
    }
}

// Preconditioned MINRES on the batch; ws.b and ws.x are set by the caller (x_nonzero: x holds an initial guess).
static int minres_batch(Solver &sv, int nsamples, bool x_nonzero)
{
    Ctx *c = sv.c;
    SaddleSys &sys = *sv.sys;
    SolveWs &ws = *sv.ws;
    const int ld = sv.ld, N = sys.N;
    const size_t NS = (size_t)N * ld;
    const Shape shN = shape_for(N, ld);
    const int sthreads = 128, sblocks = (ld + sthreads - 1) / sthreads;
    if (x_nonzero) saddle_apply(sv, EP_RESID, ws.x, ws.v1, ws.b, false, PMC_K_SADDLE_APPLY);
    else CK(cudaMemcpyAsync(ws.v1, ws.b, NS * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    fill(c, ws.v0, NS, 0.0);
    fill(c, ws.w0, NS, 0.0);
    fill(c, ws.w1, NS, 0.0);
    CK(cudaMemsetAsync(c->d_nactive, 0, sizeof(int), c->stream));
    int nblk = apply_prec(sv, ws.v1, ws.u1, true);
    launch(c, PMC_K_SCALAR, 0.0, k_minres_init, dim3(sblocks), dim3(sthreads), ld, nsamples, nblk, ws.partial, c->rel,
           c->abs_, ws.st, ws.active, ws.iters, c->d_nactive);
    CK(cudaMemcpyAsync(c->h_nactive, c->d_nactive, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    double *v0 = ws.v0, *v1 = ws.v1, *w0 = ws.w0, *w1 = ws.w1, *u1 = ws.u1, *q = ws.q;
    double *st = ws.st;
    int it = 0;
    while (*c->h_nactive > 0 && it < c->maxit) {
        const int chunk = std::min(c->check_every, c->maxit - it);
        for (int k = 0; k < chunk; ++k, ++it) {
            saddle_apply(sv, EP_AX, u1, q, nullptr, true, PMC_K_SADDLE_APPLY);
            launch(c, PMC_K_SCALAR, 0.0, k_minres_alpha, dim3(sblocks), dim3(sthreads), ld, shN.nblk, ws.partial, st,
                   ws.active, c->d_nactive);
            launch(c, PMC_K_LANCZOS_UPDATE, (double)ld * N * 4 * 8.0, k_lincomb3, shN.grid, shN.block, N, ld,
                   shN.rows_per_cta, st + (size_t)ST_CQ * ld, q, st + (size_t)ST_CV1 * ld, v1,
                   st + (size_t)ST_CV0 * ld, v0);
            nblk = apply_prec(sv, v0, q, true);
            launch(c, PMC_K_SCALAR, 0.0, k_minres_beta, dim3(sblocks), dim3(sthreads), ld, nblk, ws.partial, c->maxit,
                   st, ws.active, ws.iters, c->d_nactive);
            launch(c, PMC_K_SOLUTION_UPDATE, (double)ld * N * 6 * 8.0, k_solution_update, shN.grid, shN.block, N, ld,
                   shN.rows_per_cta, st + (size_t)ST_CW0 * ld, st + (size_t)ST_CW1 * ld, st + (size_t)ST_CU * ld,
                   st + (size_t)ST_CX * ld, w0, w1, u1, ws.x);
            std::swap(u1, q);
            std::swap(v0, v1);
            std::swap(w0, w1);
        }
        CK(cudaMemcpyAsync(c->h_nactive, c->d_nactive, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    launch(c, PMC_K_MISC, 0.0, k_sum_int, dim3(1), dim3(256), nsamples, ws.iters, c->d_iters_total);
    if (c->cuda_status != cudaSuccess)
        return fail(c, PMC_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(c->cuda_status));
    return PMC_OK;
}

// Sampler solve on the device: rhs_p = batched [Ne][ld] right-hand side at `level`; x0_p (nullable) batched initial
// guess of the Gaussian field.  The field is left in ws.x + Nf*ld.
static int sampler_solve_dev(Ctx *c, int level, int ld, int nsamples, const double *rhs_p, const double *x0_p,
                             SolveWs &ws)
{
    SaddleSys &sys = c->s[level].sys;
    Solver sv{c, &sys, &ws, nullptr, ld};
    const size_t po = (size_t)sys.Nf * ld, pn = (size_t)sys.Ne * ld;
    fill(c, ws.b, po, 0.0);  // rhs_u = 0 (/root/reference/src/PDESampler.cpp:441-442)
    CK(cudaMemcpyAsync(ws.b + po, rhs_p, pn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    fill(c, ws.x, po, 0.0);
    if (x0_p) CK(cudaMemcpyAsync(ws.x + po, x0_p, pn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    else fill(c, ws.x + po, pn, 0.0);
    return minres_batch(sv, nsamples, x0_p != nullptr);
}

// Darcy solve on the device: k_ext = batched [Ne+1][ld] (row Ne = 1).  Q_dev[ld] receives obs . sol.
static int darcy_solve_dev(Ctx *c, int level, int ld, int nsamples, const double *k_ext, SolveWs &ws, double *Q_dev)
{
    DarcyLevel &L = c->d[level];
    SaddleSys &sys = L.sys;
    Solver sv{c, &sys, &ws, k_ext, ld};
    const int N = sys.N;
    darcy_setup_values(sv);
    const Shape shN = shape_for(N, ld);
    launch(c, PMC_K_MISC, (double)ld * N * 8.0, k_broadcast, shN.grid, shN.block, N, ld, shN.rows_per_cta, nsamples,
           L.d_rhs_bc, ws.b);
    if (L.ess_nonzero) {
        // rhs_bc -= M(k)[:, ess] ess_data  (BlockMatrix::EliminateRowCol, /root/reference/src/DarcySolver.cpp:498)
        const Shape shF = shape_for(sys.Nf, ld);
        launch(c, PMC_K_MISC, (double)ld * sys.Nf * 8.0, k_broadcast, shF.grid, shF.block, sys.Nf, ld, shF.rows_per_cta,
               nsamples, L.d_ess_u_data, ws.mu_z);
        spmm(c, PMC_K_SADDLE_APPLY, EP_RESID, L.Mbc, k_ext, ld, ws.mu_z, ws.b, ws.b, nullptr, nullptr, false, 0, 0,
             nullptr, nullptr, 0, 3.0 * sys.Nf + sys.Ne);
    }
    fill(c, ws.x, (size_t)N * ld, 0.0);  // p_sol = 0 (:629)
    int rc = minres_batch(sv, nsamples, false);
    if (rc) return rc;
    if (Q_dev) {
        launch(c, PMC_K_MISC, (double)ld * N * 8.0, k_dot_fixed, shN.grid, shN.block, N, ld, shN.rows_per_cta, L.d_obs,
               ws.x, ws.partial);
        launch(c, PMC_K_MISC, 0.0, k_finish_sum, dim3((ld + 127) / 128), dim3(128), ld, shN.nblk, ws.partial, Q_dev);
    }
    return PMC_OK;
}

static int pad_ld(int nsamples)
{
    if (nsamples >= 64) return ((nsamples + 63) / 64) * 64;
    return ((nsamples + 1) / 2) * 2;
}

static int pick_batch(Ctx *c, size_t bytes_per_sample, int nsamples)
{
    int b = c->max_batch;
    if (b <= 0) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        free_b += c->arena.cap;
        const size_t usable = (size_t)(0.80 * (double)free_b);
        size_t nb = usable / std::max<size_t>(bytes_per_sample, 1);
        if (nb > 4096) nb = 4096;
        b = (int)nb;
    }
    if (b >= 64) b = (b / 64) * 64;
    if (b < 1) b = 1;
    return std::min(b, std::max(nsamples, 1));
}

static size_t solve_bytes_per_sample(const SaddleSys &sys)
{
    Arena dry;
    dry.dry = true;
    SolveWs ws;
    carve_solve(dry, sys, 64, ws);
    return dry.peak / 64 + 64;
}

static int check_level(Ctx *c, int level, bool sampler, bool darcy)
{
    if (!c) return PMC_ERR_ARG;
    if (level < 0 || level >= c->nlevels) return fail(c, PMC_ERR_ARG, "level %d out of range [0,%d)", level, c->nlevels);
    if (sampler && !c->s[level].set) return fail(c, PMC_ERR_STATE, "sampler level %d not uploaded", level);
    if (darcy && !c->d[level].set) return fail(c, PMC_ERR_STATE, "Darcy level %d not uploaded", level);
    int rc;
    if (sampler && (rc = prepare_sampler(c, level))) return rc;
    if (darcy && (rc = prepare_darcy(c, level))) return rc;
    return PMC_OK;
}

static int ensure_pinned(Ctx *c, size_t count)
{
    if (c->h_pinned_count >= count) return PMC_OK;
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    c->h_pinned = nullptr;
    c->h_pinned_count = 0;
    CK(cudaMallocHost((void **)&c->h_pinned, count * sizeof(double)));
    c->h_pinned_count = count;
    return PMC_OK;
}

static int finish(Ctx *c)
{
    CK(cudaStreamSynchronize(c->stream));
    if (c->cuda_status != cudaSuccess) {
        cudaError_t e = c->cuda_status;
        c->cuda_status = cudaSuccess;
        return fail(c, PMC_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    }
    if (c->arena.overflow) {
        c->arena.overflow = false;
        return fail(c, PMC_ERR_NOMEM, "batch workspace overflow");
    }
    return PMC_OK;
}

static int rng_launch(Ctx *c, int mode, uint64_t pos0, uint64_t pstride, uint64_t limit, int64_t nj, int64_t ni,
                      int64_t si, int64_t sj, int T, double neg_g, const double *w_sqrt, double *out, int32_t *out_i)
{
    if (!c->rng_ready) return fail(c, PMC_ERR_STATE, "pmc_rng_init has not been called");
    if (nj <= 0 || ni <= 0) return PMC_OK;
    RngArgs a;
    a.pos0 = pos0; a.pstride = pstride; a.limit = limit; a.nj = nj; a.ni = ni; a.si = si; a.sj = sj; a.T = T;
    a.mu = c->mu; a.sigma = c->sigma; a.neg_g = neg_g; a.w_sqrt = w_sqrt; a.out = out; a.out_i = out_i;
    const int64_t chunks = (ni + T - 1) / T;
    if (chunks > 65535) return fail(c, PMC_ERR_ARG, "rng: too many chunks");
    dim3 grid((unsigned)((nj + 127) / 128), (unsigned)chunks), block(128);
    const double bytes = (double)nj * (double)ni * (mode == 0 ? 4.0 : 8.0);
    if (mode == 0) launch(c, PMC_K_RNG, bytes, k_rng<0>, grid, block, a, (const RngTables *)c->d_tab);
    else if (mode == 1) launch(c, PMC_K_RNG, bytes, k_rng<1>, grid, block, a, (const RngTables *)c->d_tab);
    else launch(c, PMC_K_RNG, bytes, k_rng<2>, grid, block, a, (const RngTables *)c->d_tab);
    return PMC_OK;
}

// ==================================================================================================
// C ABI
// ==================================================================================================
extern "C" {

int pmc_create(int device, int nlevels, pmc_handle *out)
{
    if (!out || nlevels < 1) return fail(nullptr, PMC_ERR_ARG, "pmc_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, PMC_ERR_CUDA, "pmc_create: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, PMC_ERR_ARG, "pmc_create: device %d of %d", device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
        return fail(nullptr, PMC_ERR_CUDA, "pmc_create: device %d is not sm_100 class (kernels are built for sm_100a only)", device);
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, PMC_ERR_CUDA, "cudaSetDevice failed");
    Ctx *c = new Ctx();
    c->device = device;
    c->nlevels = nlevels;
    c->cfg_sampler.max_vlevels = -1;  // single-level Schur smoother when alpha*W dominates (short correlation length)
    c->cfg_darcy.omega = 2.0;         // aggregation-type coarse spaces under-correct the pressure Laplacian
    c->s.resize(nlevels);
    c->d.resize(nlevels);
    memset(&c->stats, 0, sizeof c->stats);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&c->d_nactive, sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&c->d_iters_total, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc((void **)&c->d_tab, sizeof(RngTables)) != cudaSuccess ||
        cudaMallocHost((void **)&c->h_nactive, sizeof(int)) != cudaSuccess) {
        delete c;
        return fail(nullptr, PMC_ERR_CUDA, "pmc_create: CUDA resource allocation failed");
    }
    cudaMemset(c->d_iters_total, 0, sizeof(unsigned long long));
    *out = c;
    return PMC_OK;
}

void pmc_destroy(pmc_handle c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (void *p : c->owned) cudaFree(p);
    if (c->arena.base) cudaFree(c->arena.base);
    cudaFree(c->d_nactive);
    cudaFree(c->d_iters_total);
    cudaFree(c->d_tab);
    cudaFreeHost(c->h_nactive);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (auto &ep : c->ev_pending) { cudaEventDestroy(ep.a); cudaEventDestroy(ep.b); }
    for (auto &ep : c->ev_free) { cudaEventDestroy(ep.a); cudaEventDestroy(ep.b); }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *pmc_last_error(pmc_handle c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int pmc_set_stream(pmc_handle c, void *cuda_stream)
{
    if (!c) return PMC_ERR_ARG;
    cudaStreamSynchronize(c->stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return PMC_OK;
}

int pmc_synchronize(pmc_handle c)
{
    if (!c) return PMC_ERR_ARG;
    return finish(c);
}

int pmc_set_tolerances(pmc_handle c, double rel_tol, double abs_tol, int max_iter)
{
    if (!c || max_iter < 1) return PMC_ERR_ARG;
    c->rel = rel_tol;
    c->abs_ = abs_tol;
    c->maxit = max_iter;
    return PMC_OK;
}

int pmc_set_preconditioner(pmc_handle c, int mass_degree, int schur_degree, double schur_ratio, int coarse_degree,
                           double coarse_ratio)
{
    if (!c) return PMC_ERR_ARG;
    for (PrecCfg *g : {&c->cfg_sampler, &c->cfg_darcy}) {
        if (mass_degree > 0) g->mass_degree = mass_degree;
        if (schur_degree > 0) g->schur_degree = schur_degree;
        if (schur_ratio > 1.0) g->schur_ratio = schur_ratio;
        if (coarse_degree > 0) g->coarse_degree = coarse_degree;
        if (coarse_ratio > 1.0) g->coarse_ratio = coarse_ratio;
    }
    return PMC_OK;
}

int pmc_set_option(pmc_handle c, const char *key, double value)
{
    if (!c || !key) return PMC_ERR_ARG;
    std::string k(key);
    PrecCfg *g = nullptr;
    if (k.rfind("sampler.", 0) == 0) { g = &c->cfg_sampler; k = k.substr(8); }
    else if (k.rfind("darcy.", 0) == 0) { g = &c->cfg_darcy; k = k.substr(6); }
    if (g) {
        for (int l = 0; l < c->nlevels; ++l)
            if ((g == &c->cfg_sampler ? c->s[l].sys.ready : c->d[l].sys.ready))
                return fail(c, PMC_ERR_STATE, "pmc_set_option(%s): preconditioner already built; set options before the first solve / pmc_prepare", key);
        if (k == "mass_degree" && value >= 1) g->mass_degree = (int)value;
        else if (k == "schur_degree" && value >= 1) g->schur_degree = (int)value;
        else if (k == "schur_ratio" && value > 1) g->schur_ratio = value;
        else if (k == "coarse_degree" && value >= 1) g->coarse_degree = (int)value;
        else if (k == "coarse_ratio" && value > 1) g->coarse_ratio = value;
        else if (k == "omega" && value > 0) g->omega = value;
        else if (k == "max_vlevels") g->max_vlevels = (int)value;
        else return fail(c, PMC_ERR_ARG, "pmc_set_option: bad key or value '%s' = %g", key, value);
        return PMC_OK;
    }
    if (k == "max_batch" && value >= 0) c->max_batch = (int)value;
    else if (k == "check_every" && value >= 1) c->check_every = (int)value;
    else return fail(c, PMC_ERR_ARG, "pmc_set_option: unknown key '%s'", key);
    return PMC_OK;
}

int pmc_set_batch(pmc_handle c, int max_batch, int check_every)
{
    if (!c) return PMC_ERR_ARG;
    if (max_batch >= 0) c->max_batch = max_batch;
    if (check_every > 0) c->check_every = check_every;
    return PMC_OK;
}

int pmc_upload_sampler_level(pmc_handle c, int level, int Ne, int Nf, const int *M_rowptr, const int *M_col,
                             const double *M_val, const int *B_rowptr, const int *B_col, const double *B_val,
                             const double *Wdiag, int P_cols, const int *P_rowptr, const int *P_col,
                             const double *P_val, double alpha, double matern_coeff, int lognormal)
{
    if (!c) return PMC_ERR_ARG;
    if (level < 0 || level >= c->nlevels || Ne < 1 || Nf < 1 || !M_rowptr || !B_rowptr || !Wdiag)
        return fail(c, PMC_ERR_ARG, "pmc_upload_sampler_level: bad arguments");
    CK(cudaSetDevice(c->device));
    SamplerLevel &L = c->s[level];
    if (L.set) return fail(c, PMC_ERR_STATE, "sampler level %d uploaded twice", level);
    L.Ne = Ne; L.Nf = Nf; L.alpha = alpha; L.g = matern_coeff; L.lognormal = lognormal;
    L.M = csr_copy(Nf, Nf, M_rowptr, M_col, M_val);
    L.B = csr_copy(Ne, Nf, B_rowptr, B_col, B_val);
    L.Wdiag.assign(Wdiag, Wdiag + Ne);
    std::vector<double> ws(Ne);
    for (int i = 0; i < Ne; ++i) ws[i] = std::sqrt(Wdiag[i]);  // /root/reference/src/PDESampler.cpp:248-254
    int rc = to_device(c, ws, &L.w_sqrt);
    if (rc) return rc;
    L.hasP = P_rowptr != nullptr;
    if (L.hasP) {
        L.P = csr_copy(Ne, P_cols, P_rowptr, P_col, P_val);
        if ((rc = upload_csr(c, L.P, L.dP))) return rc;
        HCsr Pt = csr_transpose(L.P);
        if ((rc = upload_csr(c, Pt, L.dPt))) return rc;
    }
    L.set = true;
    return PMC_OK;
}

int pmc_upload_darcy_level(pmc_handle c, int level, int Ne, int Nf, const int *elem_ptr, const int *elem_dofs,
                           const double *elem_mat, const int *B_rowptr, const int *B_col, const double *B_val,
                           const int *ess_u, const double *ess_data, const double *rhs, const double *obs,
                           int Pp_cols, const int *Pp_rowptr, const int *Pp_col, const double *Pp_val)
{
    if (!c) return PMC_ERR_ARG;
    if (level < 0 || level >= c->nlevels || Ne < 1 || Nf < 1 || !elem_ptr || !elem_dofs || !elem_mat || !B_rowptr ||
        !ess_u || !ess_data || !rhs || !obs)
        return fail(c, PMC_ERR_ARG, "pmc_upload_darcy_level: bad arguments");
    CK(cudaSetDevice(c->device));
    DarcyLevel &L = c->d[level];
    if (L.set) return fail(c, PMC_ERR_STATE, "Darcy level %d uploaded twice", level);
    L.Ne = Ne; L.Nf = Nf;
    L.elem_ptr.assign(elem_ptr, elem_ptr + Ne + 1);
    L.elem_dofs.assign(elem_dofs, elem_dofs + elem_ptr[Ne]);
    size_t nm = 0;
    for (int e = 0; e < Ne; ++e) {
        const size_t n = (size_t)(elem_ptr[e + 1] - elem_ptr[e]);
        nm += n * n;
    }
    for (int t = 0; t < elem_ptr[Ne]; ++t)
        if (elem_dofs[t] < 0 || elem_dofs[t] >= Nf) return fail(c, PMC_ERR_ARG, "elem_dofs out of range");
    L.elem_mat.assign(elem_mat, elem_mat + nm);
    L.B = csr_copy(Ne, Nf, B_rowptr, B_col, B_val);
    L.ess_u.assign(ess_u, ess_u + Nf);
    L.ess_data.assign(ess_data, ess_data + Nf + Ne);
    L.rhs.assign(rhs, rhs + Nf + Ne);
    L.obs.assign(obs, obs + Nf + Ne);
    L.hasP = Pp_rowptr != nullptr;
    if (L.hasP) L.Pp = csr_copy(Ne, Pp_cols, Pp_rowptr, Pp_col, Pp_val);
    L.set = true;
    return PMC_OK;
}

int pmc_prepare(pmc_handle c)
{
    if (!c) return PMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    for (int l = 0; l < c->nlevels; ++l) {
        int rc;
        if (c->s[l].set && (rc = prepare_sampler(c, l))) return rc;
        if (c->d[l].set && (rc = prepare_darcy(c, l))) return rc;
    }
    return PMC_OK;
}

// ---- RNG ------------------------------------------------------------------------------------------
int pmc_rng_init(pmc_handle c, double mu, double sigma, int nparts, int mypart)
{
    if (!c) return PMC_ERR_ARG;
    if (nparts > 1 && (mypart < 0 || mypart >= nparts)) return fail(c, PMC_ERR_ARG, "pmc_rng_init: mypart out of range");
    CK(cudaSetDevice(c->device));
    RngTables *t = new RngTables();
    h_build_rng_tables(*t, nparts, mypart);
    cudaError_t e = cudaMemcpy(c->d_tab, t, sizeof(RngTables), cudaMemcpyHostToDevice);
    delete t;
    if (e != cudaSuccess) return fail(c, PMC_ERR_CUDA, "rng table upload failed: %s", cudaGetErrorString(e));
    c->mu = mu;
    c->sigma = sigma;
    c->rng_ready = true;
    return PMC_OK;
}

static int rng_fill_common(Ctx *c, int mode, uint64_t pos, int64_t n, void *out)
{
    if (!c || n < 0 || (n > 0 && !out)) return PMC_ERR_ARG;
    if (n == 0) return PMC_OK;
    CK(cudaSetDevice(c->device));
    const size_t esz = mode == 0 ? 4 : 8;
    int rc = ensure_arena(c, (size_t)n * esz + 4096);
    if (rc) return rc;
    const int T = 64;
    const int64_t nj = (n + T - 1) / T;
    rc = rng_launch(c, mode, pos, T, (uint64_t)n, nj, T, 1, T, T, 0.0, nullptr, (double *)c->arena.base,
                    (int32_t *)c->arena.base);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, c->arena.base, (size_t)n * esz, cudaMemcpyDeviceToHost, c->stream));
    return finish(c);
}

int pmc_rng_fill_int(pmc_handle c, uint64_t pos, int64_t n, int32_t *out) { return rng_fill_common(c, 0, pos, n, out); }
int pmc_rng_fill(pmc_handle c, uint64_t pos, int64_t n, double *out) { return rng_fill_common(c, 1, pos, n, out); }

int pmc_sampler_sample_batch(pmc_handle c, int level, int nsamples, uint64_t pos0, double *xi_out)
{
    if (!c) return PMC_ERR_ARG;
    if (level < 0 || level >= c->nlevels || !c->s[level].set) return fail(c, PMC_ERR_STATE, "sampler level %d not uploaded", level);
    // realisation j occupies positions pos0 + j*Ne ... : a flat fill of nsamples*Ne values
    return rng_fill_common(c, 1, pos0, (int64_t)nsamples * c->s[level].Ne, xi_out);
}

// ---- sampler Eval ---------------------------------------------------------------------------------
// Restrict a batched right-hand side from xi_level to level (Ps^T chain, /root/reference/src/PDESampler.cpp:361-368).
// bufs: two batched buffers large enough for Ne(xi_level); returns the buffer holding the result.
static double *restrict_rhs(Ctx *c, int xi_level, int level, int ld, double *cur, double *other)
{
    for (int l = xi_level; l < level; ++l) {
        SamplerLevel &L = c->s[l];
        spmm(c, PMC_K_TRANSFER, EP_AX, L.dPt, nullptr, ld, cur, other, nullptr, nullptr, nullptr, false, 0, 0, nullptr,
             nullptr, 0, (double)L.Ne + c->s[l + 1].Ne);
        std::swap(cur, other);
    }
    return cur;
}

// Prolongate a batched Gaussian field from init_level down to level (finer) (:496-508).
static double *prolong_field(Ctx *c, int init_level, int level, int ld, double *cur, double *other)
{
    for (int l = init_level; l > level; --l) {
        SamplerLevel &L = c->s[l - 1];
        spmm(c, PMC_K_TRANSFER, EP_AX, L.dP, nullptr, ld, cur, other, nullptr, nullptr, nullptr, false, 0, 0, nullptr,
             nullptr, 0, (double)L.Ne + c->s[l].Ne);
        std::swap(cur, other);
    }
    return cur;
}

int pmc_sampler_eval_batch(pmc_handle c, int level, int xi_level, int nsamples, const double *xi, const double *init_s,
                           int init_level, int use_init, double *s_out, double *embed_s_out, int *iters_out)
{
    int rc = check_level(c, level, true, false);
    if (rc) return rc;
    if (nsamples < 0 || !xi || !s_out) return fail(c, PMC_ERR_ARG, "pmc_sampler_eval_batch: bad arguments");
    if (xi_level < 0 || xi_level > level || !c->s[xi_level].set)
        return fail(c, PMC_ERR_ARG, "xi_level %d must be an uploaded level <= level %d", xi_level, level);
    const bool warm = use_init > 0 && init_s != nullptr;
    if (warm && (init_level < level || init_level >= c->nlevels || !c->s[init_level].set))
        return fail(c, PMC_ERR_ARG, "init_level %d must be an uploaded level >= level %d", init_level, level);
    for (int l = xi_level; l < level; ++l)
        if (!c->s[l].hasP) return fail(c, PMC_ERR_STATE, "sampler level %d has no prolongator", l);
    if (warm)
        for (int l = level; l < init_level; ++l)
            if (!c->s[l].hasP) return fail(c, PMC_ERR_STATE, "sampler level %d has no prolongator", l);
    if (nsamples == 0) return PMC_OK;
    CK(cudaSetDevice(c->device));
    SamplerLevel &L = c->s[level];
    SaddleSys &sys = L.sys;
    const int Ne = L.Ne, Nex = c->s[xi_level].Ne, Nei = warm ? c->s[init_level].Ne : 0;
    const int nmax = std::max(Nex, std::max(Ne, Nei));
    const size_t per_sample = solve_bytes_per_sample(sys) + (size_t)(4 * nmax) * 8 + 64;
    const int B = pick_batch(c, per_sample, nsamples);
    const int ldB = pad_ld(B);
    if ((rc = ensure_arena(c, per_sample * (size_t)ldB + (1 << 16)))) return rc;
    for (int s0 = 0; s0 < nsamples; s0 += B) {
        const int ns = std::min(B, nsamples - s0);
        const int ld = pad_ld(ns);
        Arena &ar = c->arena;
        ar.top = 0;
        SolveWs ws;
        carve_solve(ar, sys, ld, ws);
        double *stage = ar.alloc((size_t)ns * nmax);
        double *bufA = ar.alloc((size_t)nmax * ld), *bufB = ar.alloc((size_t)nmax * ld);
        double *bufC = ar.alloc((size_t)nmax * ld);
        if (ar.overflow) return finish(c);
        // rhs_s = -g * xi * w_sqrt at xi_level (:352-358 / :423-428), then restrict
        CK(cudaMemcpyAsync(stage, xi + (size_t)s0 * Nex, (size_t)ns * Nex * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        launch(c, PMC_K_MISC, (double)ns * Nex * 16.0, k_transpose_in<1>, dim3((Nex + 31) / 32, (ld + 31) / 32),
               dim3(32, 8), Nex, ld, ns, (const double *)stage, bufA, -c->s[xi_level].g, (const double *)c->s[xi_level].w_sqrt);
        double *rhs = restrict_rhs(c, xi_level, level, ld, bufA, bufB);
        double *x0 = nullptr;
        if (warm) {
            double *t1 = (rhs == bufA) ? bufB : bufA;
            CK(cudaMemcpyAsync(stage, init_s + (size_t)s0 * Nei, (size_t)ns * Nei * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            launch(c, PMC_K_MISC, (double)ns * Nei * 16.0, k_transpose_in<0>, dim3((Nei + 31) / 32, (ld + 31) / 32),
                   dim3(32, 8), Nei, ld, ns, (const double *)stage, t1, 0.0, (const double *)nullptr);
            x0 = prolong_field(c, init_level, level, ld, t1, bufC);
        }
        if ((rc = sampler_solve_dev(c, level, ld, ns, rhs, x0, ws))) return rc;
        const double *field = ws.x + (size_t)sys.Nf * ld;
        dim3 tg((Ne + 31) / 32, (ld + 31) / 32), tb(32, 8);
        if (L.lognormal) launch(c, PMC_K_MISC, (double)ns * Ne * 16.0, k_transpose_out<1>, tg, tb, Ne, ld, ns, field, stage);
        else launch(c, PMC_K_MISC, (double)ns * Ne * 16.0, k_transpose_out<0>, tg, tb, Ne, ld, ns, field, stage);
        CK(cudaMemcpyAsync(s_out + (size_t)s0 * Ne, stage, (size_t)ns * Ne * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (embed_s_out) {
            CK(cudaStreamSynchronize(c->stream));
            launch(c, PMC_K_MISC, (double)ns * Ne * 16.0, k_transpose_out<0>, tg, tb, Ne, ld, ns, field, stage);
            CK(cudaMemcpyAsync(embed_s_out + (size_t)s0 * Ne, stage, (size_t)ns * Ne * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        if (iters_out) CK(cudaMemcpyAsync(iters_out + s0, ws.iters, (size_t)ns * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if ((rc = finish(c))) return rc;
    }
    return PMC_OK;
}

// ---- Darcy ----------------------------------------------------------------------------------------
static int darcy_host_batch(Ctx *c, int level, int nsamples, const double *k, const double *xin, double *Q_out,
                            double *C_out, double *sol_out, int *iters_out, bool apply_only)
{
    int rc = check_level(c, level, false, true);
    if (rc) return rc;
    if (nsamples < 0 || !k) return fail(c, PMC_ERR_ARG, "Darcy batch: bad arguments");
    if (nsamples == 0) return PMC_OK;
    CK(cudaSetDevice(c->device));
    DarcyLevel &L = c->d[level];
    SaddleSys &sys = L.sys;
    const int Ne = L.Ne, N = sys.N;
    const size_t per_sample = solve_bytes_per_sample(sys) + (size_t)(Ne + 1 + 2 * N) * 8 + 64;
    const int B = pick_batch(c, per_sample, nsamples);
    const int ldB = pad_ld(B);
    if ((rc = ensure_arena(c, per_sample * (size_t)ldB + (1 << 16)))) return rc;
    for (int s0 = 0; s0 < nsamples; s0 += B) {
        const int ns = std::min(B, nsamples - s0);
        const int ld = pad_ld(ns);
        Arena &ar = c->arena;
        ar.top = 0;
        SolveWs ws;
        carve_solve(ar, sys, ld, ws);
        double *stage = ar.alloc((size_t)ns * N);
        double *k_ext = ar.alloc((size_t)(Ne + 1) * ld);
        double *Qd = ar.alloc(ld);
        if (ar.overflow) return finish(c);
        CK(cudaMemcpyAsync(stage, k + (size_t)s0 * Ne, (size_t)ns * Ne * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        launch(c, PMC_K_MISC, (double)ns * Ne * 16.0, k_transpose_in<0>, dim3((Ne + 31) / 32, (ld + 31) / 32), dim3(32, 8),
               Ne, ld, ns, (const double *)stage, k_ext, 0.0, (const double *)nullptr);
        fill(c, k_ext + (size_t)Ne * ld, ld, 1.0);
        dim3 tgN((N + 31) / 32, (ld + 31) / 32), tb(32, 8);
        if (apply_only) {
            Solver sv{c, &sys, &ws, k_ext, ld};
            CK(cudaMemcpyAsync(stage, xin + (size_t)s0 * N, (size_t)ns * N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            launch(c, PMC_K_MISC, (double)ns * N * 16.0, k_transpose_in<0>, tgN, tb, N, ld, ns, (const double *)stage, ws.x,
                   0.0, (const double *)nullptr);
            saddle_apply(sv, EP_AX, ws.x, ws.q, nullptr, false, PMC_K_SADDLE_APPLY);
            launch(c, PMC_K_MISC, (double)ns * N * 16.0, k_transpose_out<0>, tgN, tb, N, ld, ns, (const double *)ws.q, stage);
            CK(cudaMemcpyAsync(sol_out + (size_t)s0 * N, stage, (size_t)ns * N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            if ((rc = finish(c))) return rc;
            continue;
        }
        if ((rc = darcy_solve_dev(c, level, ld, ns, k_ext, ws, Qd))) return rc;
        if (Q_out) CK(cudaMemcpyAsync(Q_out + s0, Qd, (size_t)ns * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (C_out)
            for (int j = 0; j < ns; ++j) C_out[s0 + j] = (double)N;  // /root/reference/src/DarcySolver.cpp:429
        if (sol_out) {
            launch(c, PMC_K_MISC, (double)ns * N * 16.0, k_transpose_out<0>, tgN, tb, N, ld, ns, (const double *)ws.x, stage);
            CK(cudaMemcpyAsync(sol_out + (size_t)s0 * N, stage, (size_t)ns * N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        if (iters_out) CK(cudaMemcpyAsync(iters_out + s0, ws.iters, (size_t)ns * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if ((rc = finish(c))) return rc;
    }
    return PMC_OK;
}

int pmc_darcy_solve_batch(pmc_handle c, int level, int nsamples, const double *k, double *Q_out, double *C_out,
                          double *sol_out, int *iters_out)
{
    if (!c) return PMC_ERR_ARG;
    return darcy_host_batch(c, level, nsamples, k, nullptr, Q_out, C_out, sol_out, iters_out, false);
}

int pmc_darcy_apply_batch(pmc_handle c, int level, int nsamples, const double *k, const double *x, double *y)
{
    if (!c) return PMC_ERR_ARG;
    if (!x || !y) return fail(c, PMC_ERR_ARG, "pmc_darcy_apply_batch: bad arguments");
    return darcy_host_batch(c, level, nsamples, k, x, nullptr, nullptr, y, nullptr, true);
}

// ---- fused manager loops --------------------------------------------------------------------------
// mode 0: MLMC level pair (or coarsest single), sums[9]; mode 1: MC single level, sums[4].
static int level_batch(Ctx *c, int level, int nlevels, int nsamples, uint64_t pos0, double *sums, double *rows,
                       int64_t *total_iters, bool mc)
{
    if (!c) return PMC_ERR_ARG;
    if (!sums || nsamples < 0) return fail(c, PMC_ERR_ARG, "level batch: bad arguments");
    if (nlevels < 1 || nlevels > c->nlevels || level < 0 || level >= nlevels)
        return fail(c, PMC_ERR_ARG, "level %d / nlevels %d out of range", level, nlevels);
    const bool coarsest = mc || (level == nlevels - 1);
    int rc = check_level(c, level, true, true);
    if (rc) return rc;
    if (!coarsest) {
        if ((rc = check_level(c, level + 1, true, true))) return rc;
        if (!c->s[level].hasP) return fail(c, PMC_ERR_STATE, "sampler level %d has no prolongator", level);
    }
    if (!c->rng_ready) return fail(c, PMC_ERR_STATE, "pmc_rng_init has not been called");
    if (nsamples == 0) return PMC_OK;
    CK(cudaSetDevice(c->device));
    SamplerLevel &SF = c->s[level];
    DarcyLevel &DF = c->d[level];
    const int Ne = SF.Ne, Nec = coarsest ? 0 : c->s[level + 1].Ne;
    const double cost = (double)DF.sys.N + (coarsest ? 0.0 : (double)c->d[level + 1].sys.N);
    size_t solve_ps = std::max(solve_bytes_per_sample(SF.sys), solve_bytes_per_sample(DF.sys));
    const size_t per_sample = solve_ps + (size_t)(2 * Ne + 2 * Nec + (Ne + 1) + 8) * 8 + 64;
    const int B = pick_batch(c, per_sample, nsamples);
    const int ldB = pad_ld(B);
    if ((rc = ensure_arena(c, per_sample * (size_t)ldB + (size_t)B * 32 + (1 << 16)))) return rc;
    if ((rc = ensure_pinned(c, 16 + (rows ? (size_t)B * 4 : 0)))) return rc;
    CK(cudaMemsetAsync(c->d_iters_total, 0, sizeof(unsigned long long), c->stream));
    for (int s0 = 0; s0 < nsamples; s0 += B) {
        const int ns = std::min(B, nsamples - s0);
        const int ld = pad_ld(ns);
        Arena &ar = c->arena;
        ar.top = 0;
        double *rhs_f = ar.alloc((size_t)Ne * ld);
        double *rhs_c = coarsest ? nullptr : ar.alloc((size_t)Nec * ld);
        double *s_c = coarsest ? nullptr : ar.alloc((size_t)Nec * ld);
        double *x0_f = coarsest ? nullptr : ar.alloc((size_t)Ne * ld);
        double *k_ext = ar.alloc((size_t)(Ne + 1) * ld);
        double *Qf = ar.alloc(ld), *Qc = ar.alloc(ld), *out9 = ar.alloc(16);
        double *rows_d = rows ? ar.alloc((size_t)ns * 4) : nullptr;
        const size_t mark = ar.top;
        if (ar.overflow) return finish(c);
        // Sample(level, xi) fused with rhs_s = -g W^{1/2} xi  (/root/reference/src/PDESampler.cpp:336-340, :352-358)
        fill(c, rhs_f, (size_t)Ne * ld, 0.0);
        if ((rc = rng_launch(c, 2, pos0 + (uint64_t)s0 * (uint64_t)Ne, (uint64_t)Ne, ~0ull, ns, Ne, ld, 1, 32, -SF.g,
                             SF.w_sqrt, rhs_f, nullptr)))
            return rc;
        if (!coarsest) {
            SamplerLevel &SC = c->s[level + 1];
            DarcyLevel &DC = c->d[level + 1];
            // Eval(level+1, xi, ., init_s, false): restrict the right-hand side, solve from zero (:431-438)
            spmm(c, PMC_K_TRANSFER, EP_AX, SF.dPt, nullptr, ld, rhs_f, rhs_c, nullptr, nullptr, nullptr, false, 0, 0,
                 nullptr, nullptr, 0, (double)Ne + Nec);
            {
                ar.top = mark;
                SolveWs ws;
                carve_solve(ar, SC.sys, ld, ws);
                if (ar.overflow) return finish(c);
                if ((rc = sampler_solve_dev(c, level + 1, ld, ns, rhs_c, nullptr, ws))) return rc;
                const double *field = ws.x + (size_t)SC.sys.Nf * ld;
                CK(cudaMemcpyAsync(s_c, field, (size_t)Nec * ld * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
                const Shape sh = shape_for(Nec, ld);
                if (SC.lognormal) launch(c, PMC_K_MISC, (double)ld * Nec * 16.0, k_map_rows<1>, sh.grid, sh.block, Nec, ld, sh.rows_per_cta, (const double *)s_c, k_ext);
                else launch(c, PMC_K_MISC, (double)ld * Nec * 16.0, k_map_rows<0>, sh.grid, sh.block, Nec, ld, sh.rows_per_cta, (const double *)s_c, k_ext);
                fill(c, k_ext + (size_t)Nec * ld, ld, 1.0);
            }
            {
                ar.top = mark;
                SolveWs ws;
                carve_solve(ar, DC.sys, ld, ws);
                if (ar.overflow) return finish(c);
                if ((rc = darcy_solve_dev(c, level + 1, ld, ns, k_ext, ws, Qc))) return rc;
            }
            // initial guess for the fine solve: prolongated coarse Gaussian field (:496-511)
            spmm(c, PMC_K_TRANSFER, EP_AX, SF.dP, nullptr, ld, s_c, x0_f, nullptr, nullptr, nullptr, false, 0, 0, nullptr,
                 nullptr, 0, (double)Ne + Nec);
        }
        {
            ar.top = mark;
            SolveWs ws;
            carve_solve(ar, SF.sys, ld, ws);
            if (ar.overflow) return finish(c);
            if ((rc = sampler_solve_dev(c, level, ld, ns, rhs_f, x0_f, ws))) return rc;
            const double *field = ws.x + (size_t)SF.sys.Nf * ld;
            const Shape sh = shape_for(Ne, ld);
            if (SF.lognormal) launch(c, PMC_K_MISC, (double)ld * Ne * 16.0, k_map_rows<1>, sh.grid, sh.block, Ne, ld, sh.rows_per_cta, field, k_ext);
            else launch(c, PMC_K_MISC, (double)ld * Ne * 16.0, k_map_rows<0>, sh.grid, sh.block, Ne, ld, sh.rows_per_cta, field, k_ext);
            fill(c, k_ext + (size_t)Ne * ld, ld, 1.0);
        }
        {
            ar.top = mark;
            SolveWs ws;
            carve_solve(ar, DF.sys, ld, ws);
            if (ar.overflow) return finish(c);
            if ((rc = darcy_solve_dev(c, level, ld, ns, k_ext, ws, Qf))) return rc;
        }
        launch(c, PMC_K_MISC, (double)ns * 16.0, k_mlmc_accumulate, dim3(1), dim3(256), ns, (const double *)Qf,
               (const double *)(coarsest ? nullptr : Qc), cost, out9, rows_d);
        CK(cudaMemcpyAsync(c->h_pinned, out9, 9 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (rows) CK(cudaMemcpyAsync(c->h_pinned + 16, rows_d, (size_t)ns * 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if ((rc = finish(c))) return rc;
        const double *o = c->h_pinned;
        if (mc) {
            // MC_Manager enum {Q2, Q, ABSQ, C} (/root/reference/src/MC_Manager.hpp:61)
            sums[0] += o[3]; sums[1] += o[4]; sums[2] += o[5]; sums[3] += o[6];
            if (rows)
                for (int j = 0; j < ns; ++j) {
                    rows[2 * (size_t)(s0 + j) + 0] = c->h_pinned[16 + 4 * j + 1];
                    rows[2 * (size_t)(s0 + j) + 1] = c->h_pinned[16 + 4 * j + 3];
                }
        } else {
            for (int k = 0; k < 9; ++k) sums[k] += o[k];
            if (rows) memcpy(rows + 4 * (size_t)s0, c->h_pinned + 16, (size_t)ns * 4 * sizeof(double));
        }
    }
    if (total_iters) {
        unsigned long long t = 0;
        CK(cudaMemcpy(&t, c->d_iters_total, sizeof t, cudaMemcpyDeviceToHost));
        *total_iters = (int64_t)t;
    }
    return PMC_OK;
}

int pmc_mlmc_level_batch(pmc_handle c, int level, int nlevels, int nsamples, uint64_t pos0, double *sums, double *rows,
                         int64_t *total_iters)
{
    return level_batch(c, level, nlevels, nsamples, pos0, sums, rows, total_iters, false);
}

int pmc_mc_level_batch(pmc_handle c, int level, int nsamples, uint64_t pos0, double *sums, double *rows,
                       int64_t *total_iters)
{
    return level_batch(c, level, c ? c->nlevels : 1, nsamples, pos0, sums, rows, total_iters, true);
}

// ---- instrumentation ------------------------------------------------------------------------------
int pmc_profile(pmc_handle c, unsigned mask)
{
    if (!c) return PMC_ERR_ARG;
    resolve_events(c);
    c->profile_mask = mask;
    return PMC_OK;
}

int pmc_reset_stats(pmc_handle c)
{
    if (!c) return PMC_ERR_ARG;
    resolve_events(c);
    memset(&c->stats, 0, sizeof c->stats);
    return PMC_OK;
}

int pmc_kernel_stats(pmc_handle c, pmc_kernel_stats_t *out)
{
    if (!c || !out) return PMC_ERR_ARG;
    resolve_events(c);
    *out = c->stats;
    return PMC_OK;
}

}  // extern "C"
