// pmc_b200.cu -- context, host-once uploads, batched solver driver and the C ABI of include/pmc_b200.h.
//
// Product path: there is no CPU fallback in this file.  Every compute entry point records a program of operations
// and runs it with the tile-persistent kernel of program.cuh (plus a few layout-conversion kernels) on the handle's
// stream, and fails with PMC_ERR_CUDA if that is not possible.
#include "../../include/pmc_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: the symbols are resolved with dlopen at the first pmc_comm_* call

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "host_sparse.hpp"
#include "program.cuh"

namespace pmc {

// --------------------------------------------------------------------------------------------------
// device containers
// --------------------------------------------------------------------------------------------------
struct DevCsr {
    int rows = 0, cols = 0, nnz = 0;
    bool weighted = false;
    // plain CSR (plain operators only; used by the per-solve set-up maps)
    int *rowptr = nullptr, *col = nullptr;
    double *val = nullptr;
    // packed sliced ELL (SLICE rows per slice, per-slice width, zero padding): the apply format (layout: program.cuh,
    // op_spmm).  The sample-independent entries of a weighted operator carry the index of a weight row that holds 1.
    int *soff = nullptr;
    unsigned char *spk = nullptr;
    int max_width = 0;  // widest slice
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
    int schur_degree_coarse = -1; // ... on the V-levels between the finest and the coarsest one: 0 = the same as schur_degree,
                                  // -1 = choose (Darcy with the hierarchy's own coarse spaces: 3 -- those levels cost little and
                                  // the better coarse correction saves 1.5 % of the iterations, bench level-0 batch 32.8 ->
                                  // 31.8 ms; aggregation coarse spaces (SPE10 geometry) and the sampler: as schur_degree)
    double schur_ratio = 4.0;  // smoother targets the eigenvalues in [1/ratio, 1] of the l1-scaled operator
    int coarse_degree = 8;     // Chebyshev steps on the coarsest V-cycle level
    double coarse_ratio = 30.0;
    bool coarse_user = false;  // set through the API: keep them.  Otherwise Darcy with the hierarchy's own coarse spaces runs
                               // degree 6 over [1/20, 1]: the coarsest level was over-solved (bench level batches 31.75 / 12.07 /
                               // 1.75 -> 31.60 / 11.76 / 1.56 ms at 0.5 % more iterations)
    double omega = 1.0;        // over-correction factor of the piecewise-constant coarse-grid correction
    double cheb_lo_scale = 0.96, cheb_hi_scale = 1.01;  // sampler, Chebyshev semi-iteration: safety margins applied to the Ritz
                               // estimates of the extreme eigenvalues of diag(H)^-1 H (a too narrow interval is caught by the
                               // residual check after the a-priori steps and costs restarted blocks; tests use that)
    double mass_scale = 1.0;   // relative scaling of the two preconditioner blocks: the Jacobi mass block is (mass_scale * theta * D)^-1
                               // (MINRES is invariant under a common factor, so one parameter covers both blocks; degree 1 only)
    bool omega_user = false;   // set through pmc_set_option: keep it whatever coarse spaces are chosen
    double p_smooth = 0.0;     // hierarchy coarse spaces: damping of one Jacobi step (on the operator at k = 1) applied to the
                               // hierarchy's piecewise-constant L2 prolongators at set-up; 0 = use them as uploaded
    int amg_passes = 0;        // aggregation coarse spaces only: pairwise matching passes per level (aggregates of up to
                               // 2^passes rows); 0 = default (3)
    double amg_smooth = -1.0;  // aggregation coarse spaces only: damping of one Jacobi smoothing step applied to the tentative
                               // prolongators at set-up (smoothed aggregation on the strength-filtered operator at k = 1,
                               // fixed across realisations); 0 = plain aggregation, < 0 = default (0.9)
    int max_vlevels = 0;       // 0: as deep as the hierarchy allows; -1 (sampler): decide from the mass term
    int method = -1;           // sampler only: 0 = MINRES on the saddle system, 1 = Jacobi-PCG on its SPD form (u eliminated
                               // system (M + alpha^-1 B^T W^-1 B) u = alpha^-1 B^T W^-1 f), 2 = Chebyshev semi-iteration on
                               // the same SPD form, -1 = the SPD form when alpha W dominates (PMC_SAMPLER_AUTO picks 1 or 2)
    int amg = -1;              // Schur V-cycle coarse spaces: 0 the hierarchy's own L2 prolongators, 1 strength-aware
                               // pairwise aggregation built here, -1 choose (aggregation when the couplings are anisotropic)
};

struct SaddleSys {
    bool ready = false, weighted = false;
    PrecCfg cfg;               // frozen at prepare time
    int Nf = 0, Ne = 0, N = 0;
    DevCsr A;                       // block operator over N rows
    DevCsr Muu;                     // RT mass block (Nf rows)
    double *dinvM_fixed = nullptr;  // sampler: 1/diag(M)
    DevCsr Dm;                      // Darcy: diag M(k) = Dm * k_ext   (Nf x (Ne+1))
    double m_lo = 0.5, m_hi = 1.5;  // spectrum of diag(M)^-1 M
    std::vector<VLevel> v;
    std::vector<HCsr> own_P;        // aggregation prolongators built here (cfg.amg)
    double *d_coarse_coef = nullptr;  // {ca_j, cb_j} of the coarsest level's Chebyshev iteration (OP_CHEB_SMALL)
    // sampler, SPD form (emit_sampler_pcg): Bs = diag(1/(alpha W)) B (Ne x Nf), MBt = [M | B^T] (Nf x N),
    // dinvH = 1/diag(M + alpha^-1 B^T W^-1 B), inv_aw = 1/(alpha W)
    // Darcy: the block operator split by row block for the apply inside the Krylov loop: Au = [M(k) | B^T] (Nf x N,
    // weighted) and Bp = B (Ne x Nf, plain: no weight gathers on the pressure rows)
    bool split_apply = false;
    DevCsr Au, Bp;
    bool pcg = false;
    DevCsr Bs, MBt;
    double *dinvH = nullptr, *inv_aw = nullptr;
    // spectrum of diag(H)^-1 H (Lanczos at set-up, widened by a safety margin): the Chebyshev semi-iteration of
    // emit_sampler_cheb runs on [h_lo, h_hi]
    bool cheb = false;
    double h_lo = 0.0, h_hi = 0.0;
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
    // optional transfer of the sampled field to the forward problem's mesh (row scale folded into the values)
    bool hasT = false;
    int n_out = 0;
    HCsr T;
    DevCsr dT;
    int out_size() const { return hasT ? n_out : Ne; }
};

struct DarcyLevel {
    bool set = false, hasP = false, ess_nonzero = false;
    int Ne = 0, Nf = 0;
    std::vector<int> elem_ptr, elem_dofs, ess_u;
    std::vector<double> elem_mat, ess_data, rhs, obs;
    HCsr B, Pp;
    double *d_rhs_bc = nullptr, *d_obs = nullptr, *d_ess_u_data = nullptr;
    int obs_nnz = 0;          // non-zero entries of obs, and their (row, value) list for the functional-only solves
    int *d_obs_idx = nullptr;
    double *d_obs_val = nullptr;
    int *d_rowmap = nullptr;  // caller's numbering of the N unknowns -> the library's (RT dofs renumbered for locality)
    std::vector<int> h_rowmap;
    // Bayesian inverse problem: m normalised pressure functionals [m][Ne], observed data and noise variance
    int n_obs = 0;
    double noise = 0.0;
    std::vector<double> h_gobs_func, h_Gobs;
    double *d_gobs_func = nullptr, *d_Gobs = nullptr;
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
};

// Device buffers of the uploaded (immutable) operators.  A handle and its clones (pmc_clone) share one store: the
// operators of a hierarchy exist once per device however many handles (one per level in the managers) run on it.
struct DevStore {
    int device = 0;
    std::mutex mu;
    std::vector<void *> owned;
    ~DevStore()
    {
        cudaSetDevice(device);
        for (void *p : owned) cudaFree(p);
    }
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
    int max_batch = 0, force_nt = 0, force_cs = 0;
    int force_group = 0;  // grid-group size G (option "group_size"; 0: choose, -1: never)
    int solo_rows = 512;  // grid-group mode: operations with at most this many rows run on the group's first CTA alone
    void *d_grp = nullptr;   // grid-group barrier words and dot-product shares
    size_t grp_cap = 0;
    bool staging = true;  // stage operator entries through shared memory (option "stage_operators")
    bool defer_x = true;  // option "defer_x"
    bool cheb3 = true;    // option "cheb_three_term": sampler Chebyshev steps without a separate update vector
    bool qoi_only = true; // option "qoi_only": Darcy solves that return Q only never form the solution vector
    bool fuse_coarse = true;  // option "fuse_coarse"
    bool stage_wide = true;   // option "stage_wide": slices wider than the staging buffers are staged chunk by chunk
    bool split_apply = true;  // option "split_apply": Darcy operator applied per row block (RT rows weighted, pressure rows plain)
    bool renumber = true;  // option "renumber": first-touch renumbering of the RT dofs inside the library
    bool single_wave = false;  // option "single_wave": prefer one wave of smaller CTAs over a mostly empty second wave
    std::vector<SamplerLevel> s;
    std::vector<DarcyLevel> d;
    // rng
    bool rng_ready = false;
    double mu = 0.0, sigma = 1.0;
    int rng_nparts = 1, rng_mypart = 0;
    RngTables *d_tab = nullptr;
    // memory
    Arena arena;
    std::shared_ptr<DevStore> store;
    int num_sms = 148;   // multiProcessorCount of the handle's device
    // device-resident copies of the most recent per-sample results (option "cache_results"): the noise of
    // pmc_sampler_sample_batch, and s_out / embed_s_out of pmc_sampler_eval_batch, in the host layout [nsamples][n]; a
    // following call that passes NULL for the corresponding input reads them instead of a host buffer
    struct ResultCache { double *p = nullptr; size_t cap = 0; int level = -1, ns = 0, n = 0; bool valid = false; };
    bool cache_results = false;
    ResultCache cache[3];
    // NCCL communicator over the ranks that share the sample budget (pmc_comm_init); null for a single rank
    void *nccl_comm = nullptr;
    int comm_ranks = 1, comm_rank = 0;
    double *d_comm_buf = nullptr;
    size_t comm_buf_count = 0;
    Op *d_ops = nullptr, *h_ops = nullptr;  // program buffer (device / pinned staging)
    size_t ops_cap = 0;
    ProgStats *d_pstats = nullptr;
    double *h_pinned = nullptr;  // pinned scratch for small results
    size_t h_pinned_count = 0;
    // stats
    pmc_kernel_stats_t stats;
    double kernel_ms = 0.0;
    int64_t kernel_launches = 0, other_launches = 0;
    unsigned long long iters_seen = 0;  // host mirror of the device iteration counter
    std::vector<EventPair> ev_pending;
    std::vector<EventPair> ev_free;
    cudaError_t cuda_status = cudaSuccess;
};

typedef pmc_context_s Ctx;

static thread_local std::string g_create_error;   // message of the last failed call that has no handle
static std::vector<double> cheb_coefficients(double lo, double hi, int deg);

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
// launch helper (the few kernels outside the persistent program: layout conversion, RNG fills, moment sums)
// --------------------------------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
static void launch(Ctx *c, int kclass, double bytes, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args)
{
    kernel<<<grid, block, 0, c->stream>>>(args...);
    c->stats.launches[kclass]++;
    c->stats.algo_bytes[kclass] += bytes;
    c->other_launches++;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess && c->cuda_status == cudaSuccess) c->cuda_status = e;
}

// --------------------------------------------------------------------------------------------------
// uploads
// --------------------------------------------------------------------------------------------------
// The RT (face) dofs are renumbered inside the library in order of first appearance when the elements are walked in
// their own order, so that the faces of an element, its pressure dof and its neighbours' sit at proportional positions
// of the u and p blocks: the gathers of the saddle apply then stay in a band that lives in L1, whatever numbering the
// mesh generator chose (direction-blocked Cartesian numberings stream the p block three times per apply).  The block
// structure [u; p] is kept; solutions and operands cross the ABI in the caller's numbering (k_to_tiles / k_from_tiles).
static int renumber_block()
{
    const char *e = getenv("PMC_RENUMBER_BLOCK");  // diagnostic override
    const int b = e ? atoi(e) : 4 * SLICE;  // measured on the bench hierarchy: 1 / 16 / 64 / 256 / all = 46.5 / 46.5 / 46.2 / 46.2 / 47.2 ms
    return b > 0 ? b : 1;
}
static std::vector<int> first_touch_order(int Nf, int Ne, const int *ptr, const int *dofs)
{
    // Elements are walked in blocks of 64; inside a block the dofs are taken local slot by local slot (all first
    // faces of the block's elements, then all second faces, ...).  A slice of 16 consecutive rows then mostly holds
    // the same kind of face of 16 consecutive elements, so the k-th gathers of its rows fall into a few contiguous
    // lines (coalesced), while the block keeps everything within a band (local).
    std::vector<int> perm(Nf, -1);
    int next = 0;
    const int blk = renumber_block();
    for (int e0 = 0; e0 < Ne; e0 += blk) {
        const int e1 = std::min(Ne, e0 + blk);
        int slots = 0;
        for (int e = e0; e < e1; ++e) slots = std::max(slots, ptr[e + 1] - ptr[e]);
        for (int sl = 0; sl < slots; ++sl)
            for (int e = e0; e < e1; ++e)
                if (ptr[e] + sl < ptr[e + 1] && perm[dofs[ptr[e] + sl]] < 0) perm[dofs[ptr[e] + sl]] = next++;
    }
    for (int f = 0; f < Nf; ++f)
        if (perm[f] < 0) perm[f] = next++;
    return perm;
}
static HCsr csr_permuted(const HCsr &A, const int *rowperm, const int *colperm)
{
    std::vector<Coo> e;
    e.reserve(A.col.size());
    for (int i = 0; i < A.rows; ++i)
        for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p)
            e.push_back({rowperm ? rowperm[i] : i, colperm ? colperm[A.col[p]] : A.col[p], A.val[p]});
    return csr_from_coo(A.rows, A.cols, e);
}
template <typename T>
static void permute_head(std::vector<T> &v, const std::vector<int> &perm)
{
    std::vector<T> t(v.begin(), v.begin() + perm.size());
    for (size_t i = 0; i < perm.size(); ++i) v[perm[i]] = t[i];
}

template <typename T>
static int to_device(Ctx *c, const std::vector<T> &h, T **out)
{
    *out = nullptr;
    const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    void *p = nullptr;
    CK(cudaMalloc(&p, bytes));
    {
        std::lock_guard<std::mutex> lk(c->store->mu);
        c->store->owned.push_back(p);
    }
    if (!h.empty()) CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<T *>(p);
    return PMC_OK;
}

// Packed sliced-ELL conversion of rows given as [begin, end) ranges into (col, val, widx) arrays.
struct HSell {
    std::vector<int> off;
    std::vector<unsigned char> pk;
    int max_width = 0;
};
static HSell make_sell(int rows, const std::vector<int> &beg, const std::vector<int> &end, const std::vector<int> &col,
                       const std::vector<double> &val, const std::vector<int> *widx)
{
    HSell S;
    const int nsl = (rows + SLICE - 1) / SLICE;
    const size_t es = widx ? 16 : 12;
    S.off.assign(nsl + 1, 0);
    for (int sl = 0; sl < nsl; ++sl) {
        int w = 0;
        for (int r = sl * SLICE; r < std::min(rows, (sl + 1) * SLICE); ++r) w = std::max(w, end[r] - beg[r]);
        S.off[sl + 1] = S.off[sl] + w;
        S.max_width = std::max(S.max_width, w);
    }
    S.pk.assign(std::max<size_t>((size_t)S.off[nsl] * SLICE * es, 16), 0);
    for (int sl = 0; sl < nsl; ++sl) {
        const size_t w = (size_t)(S.off[sl + 1] - S.off[sl]);
        unsigned char *base = S.pk.data() + (size_t)S.off[sl] * SLICE * es;
        double *pv = reinterpret_cast<double *>(base);
        int *pc = reinterpret_cast<int *>(base + w * SLICE * 8);
        int *pw = pc + w * SLICE;
        for (int r = sl * SLICE; r < std::min(rows, (sl + 1) * SLICE); ++r) {
            const int rs = r % SLICE;
            for (int p = beg[r], k = 0; p < end[r]; ++p, ++k) {
                pv[(size_t)k * SLICE + rs] = val[p];
                pc[(size_t)k * SLICE + rs] = col[p];
                if (widx) pw[(size_t)k * SLICE + rs] = (*widx)[p];
            }
        }
    }
    return S;
}

static int upload_sell(Ctx *c, const HSell &S, DevCsr &D)
{
    if (const char *pre = getenv("PMC_DUMP_SELL")) {  // diagnostic: packed operators for the stand-alone micro-benchmarks
        static int seq = 0;
        char name[512];
        snprintf(name, sizeof name, "%s_%03d_r%d_c%d_%s.bin", pre, seq++, D.rows, D.cols, D.weighted ? "w" : "p");
        if (FILE *f = fopen(name, "wb")) {
            const int hdr[4] = {D.rows, D.cols, D.weighted ? 1 : 0, (int)S.off.size()};
            const long long nb = (long long)S.pk.size();
            fwrite(hdr, sizeof hdr, 1, f);
            fwrite(&nb, sizeof nb, 1, f);
            fwrite(S.off.data(), sizeof(int), S.off.size(), f);
            fwrite(S.pk.data(), 1, S.pk.size(), f);
            fclose(f);
        }
    }
    int rc;
    if ((rc = to_device(c, S.off, &D.soff))) return rc;
    if ((rc = to_device(c, S.pk, &D.spk))) return rc;
    D.max_width = S.max_width;
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
    std::vector<int> beg(A.rowptr.begin(), A.rowptr.end() - 1), end(A.rowptr.begin() + 1, A.rowptr.end());
    return upload_sell(c, make_sell(A.rows, beg, end, A.col, A.val, nullptr), D);
}

// const_widx: the weight row that holds the constant 1 (for the operator's sample-independent entries)
static int upload_wcsr(Ctx *c, const HWCsr &A, DevCsr &D, int const_widx)
{
    D.rows = A.rows;
    D.cols = A.cols;
    D.nnz = (int)A.col.size();
    D.weighted = true;
    std::vector<int> beg(A.rows), end(A.rows), widx(A.widx);
    for (int r = 0; r < A.rows; ++r) {
        beg[r] = A.rowptr2[2 * r];
        end[r] = A.rowptr2[2 * r + 2];
        for (int p = A.rowptr2[2 * r + 1]; p < A.rowptr2[2 * r + 2]; ++p) {
            if (const_widx < 0) return fail(c, PMC_ERR_STATE, "weighted operator with fixed entries but no constant weight row");
            widx[p] = const_widx;
        }
    }
    return upload_sell(c, make_sell(A.rows, beg, end, A.col, A.val, &widx), D);
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

// ---- strength-aware aggregation for the Schur complement (host, once) --------------------------------------
// One pass of pairwise matching: every unmatched node is paired with its most strongly (negatively) coupled unmatched
// neighbour if that coupling is at least theta times its strongest coupling; otherwise it stays alone.
static HCsr pairwise_aggregate(const HCsr &S, double theta)
{
    const int n = S.rows;
    std::vector<int> agg(n, -1);
    int nc = 0;
    for (int i = 0; i < n; ++i) {
        if (agg[i] >= 0) continue;
        double smax = 0.0, best = 0.0;
        int jb = -1;
        for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p) {
            const int j = S.col[p];
            if (j == i || S.val[p] >= 0.0) continue;
            smax = std::max(smax, -S.val[p]);
            if (agg[j] < 0 && -S.val[p] > best) { best = -S.val[p]; jb = j; }
        }
        agg[i] = nc;
        if (jb >= 0 && best >= theta * smax) agg[jb] = nc;
        ++nc;
    }
    HCsr P;
    P.rows = n;
    P.cols = nc;
    P.rowptr.resize(n + 1);
    P.col.resize(n);
    P.val.assign(n, 1.0);
    for (int i = 0; i <= n; ++i) P.rowptr[i] = i;
    for (int i = 0; i < n; ++i) P.col[i] = agg[i];
    return P;
}

// Aggregation chain: two pairwise passes per level (aggregates of up to four strongly coupled nodes: semi-coarsening
// along the strong direction of anisotropic operators), Galerkin coarse operators, until the level is small.
// P <- (I - omega D_f^-1 S_f) P with S_f the strength-filtered operator (couplings weaker than theta times the row's
// strongest are lumped onto the diagonal, so the smoothing widens P along the strong direction only and the Galerkin
// operators stay sparse) -- smoothed aggregation with a prolongator that is fixed across realisations.
static HCsr smooth_prolongator(const HCsr &S, const HCsr &P, double omega, double theta)
{
    const int n = S.rows;
    HCsr Sf;
    Sf.rows = Sf.cols = n;
    Sf.rowptr.assign(n + 1, 0);
    std::vector<double> diag(n, 0.0);
    for (int i = 0; i < n; ++i) {
        double smax = 0.0, d = 0.0;
        for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p)
            if (S.col[p] != i) smax = std::max(smax, -S.val[p]);
        for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p) {
            const int j = S.col[p];
            if (j == i) d += S.val[p];
            else if (-S.val[p] >= theta * smax && smax > 0.0) { Sf.col.push_back(j); Sf.val.push_back(S.val[p]); }
            else d += S.val[p];   // weak coupling: lumped
        }
        Sf.col.push_back(i);
        Sf.val.push_back(d);
        diag[i] = d;
        Sf.rowptr[i + 1] = (int)Sf.col.size();
    }
    HCsr SP = csr_matmul(Sf, P);
    std::vector<Coo> e;
    e.reserve(SP.col.size() + P.col.size());
    for (int i = 0; i < n; ++i) {
        for (int p = P.rowptr[i]; p < P.rowptr[i + 1]; ++p) e.push_back({i, P.col[p], P.val[p]});
        const double w = diag[i] > 0.0 ? omega / diag[i] : 0.0;
        for (int p = SP.rowptr[i]; p < SP.rowptr[i + 1]; ++p) e.push_back({i, SP.col[p], -w * SP.val[p]});
    }
    return csr_from_coo(n, P.cols, e);
}

static std::vector<HCsr> aggregation_chain(HCsr S, int min_size, int max_levels, double smooth = 0.0, int passes = 2)
{
    std::vector<HCsr> Ps;
    while (S.rows > min_size && (int)Ps.size() < max_levels) {
        HCsr P = pairwise_aggregate(S, 0.25);
        for (int pass = 1; pass < passes; ++pass) {
            HCsr Sp = csr_matmul(csr_transpose(P), csr_matmul(S, P));
            P = csr_matmul(P, pairwise_aggregate(Sp, 0.25));
        }
        if (P.cols > 0.8 * S.rows) break;
        if (smooth > 0.0) P = smooth_prolongator(S, P, smooth, 0.25);
        S = csr_matmul(csr_transpose(P), csr_matmul(S, P));
        Ps.push_back(std::move(P));
    }
    return Ps;
}

// More than half of the rows have off-diagonal couplings that differ by more than a factor 4.
static bool couplings_anisotropic(const HCsr &S)
{
    int aniso = 0, counted = 0;
    for (int i = 0; i < S.rows; ++i) {
        double lo = 1e300, hi = 0.0;
        for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p)
            if (S.col[p] != i && S.val[p] != 0.0) {
                lo = std::min(lo, std::fabs(S.val[p]));
                hi = std::max(hi, std::fabs(S.val[p]));
            }
        if (hi > 0.0) { ++counted; if (hi > 4.0 * lo) ++aniso; }
    }
    return counted > 0 && 2 * aniso > counted;
}

// ---- sampler system (everything fixed across samples) ---------------------------------------------------
// sampler.method = -1 and the SPD form applies: Chebyshev semi-iteration (default) or PCG (PMC_SAMPLER_AUTO=pcg)
static bool sampler_auto_cheb()
{
    const char *e = getenv("PMC_SAMPLER_AUTO");
    return !(e && std::string(e) == "pcg");
}

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
    {
        // SPD form of the sampler system (the elimination /root/reference/src/PDESampler_Legacy.cpp:172-176,284-323 makes):
        // s = (B u - f) / (alpha W) turns [M B^T; B -alpha W] [u; s] = [0; f] into
        //   H u = alpha^-1 B^T W^-1 f,   H = M + alpha^-1 B^T W^-1 B   (symmetric positive definite).
        // When alpha W dominates the Schur complement (short correlation length) H is well conditioned under Jacobi
        // scaling and CG on its Nf rows moves ~40 % fewer bytes per iteration than MINRES on the N-row saddle system,
        // at the same iteration count; otherwise (long correlation lengths: grad-div dominated) MINRES with the Schur
        // V-cycle stays the method.
        double lmin = 1.0;
        for (int i = 0; i < Ne; ++i) {
            double t = 0;
            for (int p = S.rowptr[i]; p < S.rowptr[i + 1]; ++p) t += std::fabs(S.val[p]);
            if (t > 0) lmin = std::min(lmin, L.alpha * L.Wdiag[i] / t);
        }
        sys.pcg = sys.cfg.method == 1 || sys.cfg.method == 2 || (sys.cfg.method < 0 && lmin >= 1.0 / 20.0);
        sys.cheb = sys.pcg && (sys.cfg.method == 2 || (sys.cfg.method < 0 && sampler_auto_cheb()));
        if (sys.pcg) {
            std::vector<double> iaw(Ne);
            for (int i = 0; i < Ne; ++i) iaw[i] = 1.0 / (L.alpha * L.Wdiag[i]);
            HCsr Bsc = L.B;
            for (int i = 0; i < Ne; ++i)
                for (int p = Bsc.rowptr[i]; p < Bsc.rowptr[i + 1]; ++p) Bsc.val[p] *= iaw[i];
            std::vector<Coo> e;
            e.reserve(L.M.nnz() + L.B.nnz());
            std::vector<double> hd(Nf, 0.0);
            for (int i = 0; i < Nf; ++i) {
                for (int p = L.M.rowptr[i]; p < L.M.rowptr[i + 1]; ++p) {
                    e.push_back({i, L.M.col[p], L.M.val[p]});
                    if (L.M.col[p] == i) hd[i] += L.M.val[p];
                }
                for (int p = Bt.rowptr[i]; p < Bt.rowptr[i + 1]; ++p) {
                    e.push_back({i, Nf + Bt.col[p], Bt.val[p]});
                    hd[i] += Bt.val[p] * Bt.val[p] * iaw[Bt.col[p]];
                }
            }
            HCsr MBt = csr_from_coo(Nf, N, e);
            for (int i = 0; i < Nf; ++i) hd[i] = hd[i] != 0.0 ? 1.0 / hd[i] : 1.0;
            if ((rc = upload_csr(c, Bsc, sys.Bs))) return rc;
            if ((rc = upload_csr(c, MBt, sys.MBt))) return rc;
            if ((rc = to_device(c, hd, &sys.dinvH))) return rc;
            if ((rc = to_device(c, iaw, &sys.inv_aw))) return rc;
            if (sys.cheb) {
                // extreme eigenvalues of D^-1/2 H D^-1/2, D = diag(H) (hd holds 1/diag): H is the same for every realisation,
                // so the number of Chebyshev steps for a given residual reduction is known before the first solve
                std::vector<double> ds(Nf), t1(Nf), t2(Ne);
                for (int i = 0; i < Nf; ++i) ds[i] = std::sqrt(hd[i]);
                auto applyH = [&](const double *x, double *y) {
                    for (int i = 0; i < Nf; ++i) t1[i] = ds[i] * x[i];
                    csr_mult(L.M, t1.data(), y);
                    csr_mult(L.B, t1.data(), t2.data());
                    for (int i = 0; i < Ne; ++i) t2[i] *= iaw[i];
                    for (int i = 0; i < Nf; ++i) {
                        double a = 0;
                        for (int p = Bt.rowptr[i]; p < Bt.rowptr[i + 1]; ++p) a += Bt.val[p] * t2[Bt.col[p]];
                        y[i] = ds[i] * (y[i] + a);
                    }
                };
                double lo = 1.0, hi = 1.0;
                lanczos_extremes(Nf, 80, applyH, &lo, &hi);
                sys.h_hi = sys.cfg.cheb_hi_scale * hi;
                sys.h_lo = sys.cfg.cheb_lo_scale * lo;
                if (getenv("PMC_DEBUG_CHEB")) fprintf(stderr, "[pmc] sampler level %d: spectrum of D^-1 H in [%.5f, %.5f] (Ritz), using [%.5f, %.5f]\n", level, lo, hi, sys.h_lo, sys.h_hi);
                if (!(sys.h_lo > 0.0) || !(sys.h_hi > sys.h_lo)) sys.cheb = false;
                // Self-test of the interval, once, on the host: the Chebyshev iteration the device will run, applied to a
                // pseudo-random right-hand side, must deliver the reduction its step count promises.  An upper end below
                // the true lambda_max (unconverged Lanczos) would make the iteration diverge on the device; in that case the
                // level falls back to PCG, which needs no spectrum.  Skipped when the margins were set by hand.
                if (sys.cheb && sys.cfg.cheb_lo_scale == 0.96 && sys.cfg.cheb_hi_scale == 1.01) {
                    const double target = 1e-6;
                    const double th = 0.5 * (sys.h_hi + sys.h_lo), de = 0.5 * (sys.h_hi - sys.h_lo), sg = th / de;
                    const int m = std::max(2, (int)std::ceil(std::acosh(1.0 / target) / std::acosh(sg)));
                    std::vector<double> cf = cheb_coefficients(sys.h_lo, sys.h_hi, m);
                    std::vector<double> b(Nf), z(Nf, 0.0), d(Nf, 0.0), r(Nf);
                    uint64_t sd = 0x2545F4914F6CDD1DULL;
                    for (int i = 0; i < Nf; ++i) {
                        sd = sd * 6364136223846793005ULL + 1442695040888963407ULL;
                        b[i] = (double)(sd >> 40) / (double)(1 << 24) - 0.5;
                    }
                    // scaled system: A = D^-1/2 H D^-1/2 (applyH), plain Chebyshev (the device's D^-1 form is similar to it)
                    double n0 = 0, n1 = 0;
                    for (int i = 0; i < Nf; ++i) n0 += b[i] * b[i];
                    for (int j = 0; j < m; ++j) {
                        applyH(z.data(), r.data());
                        for (int i = 0; i < Nf; ++i) {
                            d[i] = cf[2 * j] * d[i] + cf[2 * j + 1] * (b[i] - r[i]);
                            z[i] += d[i];
                        }
                    }
                    applyH(z.data(), r.data());
                    for (int i = 0; i < Nf; ++i) n1 += (b[i] - r[i]) * (b[i] - r[i]);
                    const double red = std::sqrt(n1 / std::max(n0, 1e-300));
                    if (getenv("PMC_DEBUG_CHEB")) fprintf(stderr, "[pmc] sampler level %d: Chebyshev self-test, %d steps: reduction %.3e (target %.0e)\n", level, m, red, target);
                    if (!(red <= 1.5 * target)) sys.cheb = false;
                }
            }
        }
    }
    if (sys.cfg.max_vlevels != 1 && (sys.cfg.amg == 1 || (sys.cfg.amg < 0 && couplings_anisotropic(S)))) {
        if (sys.cfg.amg_smooth < 0.0) sys.cfg.amg_smooth = 0.9;   // sampler at half-scale SPE10: 80 -> 67 its, 36.7 -> 29.7 ms
        if (sys.cfg.amg_passes <= 0) sys.cfg.amg_passes = 3;
        sys.own_P = aggregation_chain(S, 64, 16, sys.cfg.amg_smooth, sys.cfg.amg_passes);
        Ps.clear();
        for (const HCsr &P : sys.own_P) Ps.push_back(&P);
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
    {
        std::vector<double> cc = cheb_coefficients(1.0 / sys.cfg.coarse_ratio, 1.0, sys.cfg.coarse_degree);
        if ((rc = to_device(c, cc, &sys.d_coarse_coef))) return rc;
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
        {   // rows [0, Nf) of A as their own weighted operator, rows [Nf, N) = B as a plain one
            std::vector<WEntry> eU;
            for (const WEntry &t : eA)
                if (t.r < Nf) eU.push_back(t);
            HWCsr Au = wcsr_from_entries(Nf, N, eU);
            if ((rc = upload_wcsr(c, Au, sys.Au, Ne))) return rc;
            if ((rc = upload_csr(c, Be, sys.Bp))) return rc;
            sys.split_apply = true;
        }
        HWCsr A = wcsr_from_entries(N, N, eA);
        if ((rc = upload_wcsr(c, A, sys.A, Ne))) return rc;
        HWCsr M = wcsr_from_entries(Nf, Nf, eM);
        if ((rc = upload_wcsr(c, M, sys.Muu, Ne))) return rc;
        HWCsr Mbc = wcsr_from_entries(Nf, Nf, eBc);
        if ((rc = upload_wcsr(c, Mbc, L.Mbc, Ne))) return rc;
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
        std::vector<int> oi;
        std::vector<double> ov;
        for (size_t i = 0; i < L.obs.size(); ++i)
            if (L.obs[i] != 0.0) { oi.push_back((int)i); ov.push_back(L.obs[i]); }
        L.obs_nnz = (int)oi.size();
        if ((rc = to_device(c, oi, &L.d_obs_idx))) return rc;
        if ((rc = to_device(c, ov, &L.d_obs_val))) return rc;
    }
    // Schur complement S(k) = Be diag(M(k))^-1 Be^T: pattern, unique-value map T_0 and the Galerkin chain
    std::vector<const HCsr *> Ps;
    for (int m = level; m < c->nlevels - 1 && c->d[m].set && c->d[m].hasP; ++m) Ps.push_back(&c->d[m].Pp);
    if (sys.cfg.max_vlevels != 1 && sys.cfg.amg != 0) {
        // S at k = 1 decides the aggregates (fixed for all realisations; the values stay per sample)
        std::vector<double> Md(Nf, 0.0);
        {
            size_t off = 0;
            for (int el = 0; el < Ne; ++el) {
                const int n = L.elem_ptr[el + 1] - L.elem_ptr[el];
                const int *dof = L.elem_dofs.data() + L.elem_ptr[el];
                for (int a = 0; a < n; ++a) Md[dof[a]] += L.elem_mat[off + (size_t)a * n + a];
                off += (size_t)n * n;
            }
        }
        HCsr Bs = Be;
        for (int i = 0; i < Ne; ++i)
            for (int p = Bs.rowptr[i]; p < Bs.rowptr[i + 1]; ++p) Bs.val[p] /= (Md[Bs.col[p]] > 0 ? Md[Bs.col[p]] : 1.0);
        HCsr S1 = csr_matmul(Bs, Bet);
        if (sys.cfg.amg == 1 || couplings_anisotropic(S1)) {
            // Smoothed aggregation: aggregates of up to 8 rows along the strong couplings (three pairwise passes), tentative
            // prolongators smoothed by one damped Jacobi step on the strength-filtered operator at k = 1 (fixed across
            // realisations; the Galerkin values stay per sample).  Measured on the SPE10 geometry at half scale
            // (tools/spe10_prec_sweep.py), Darcy iterations per solve / ms for 8 realisations: plain aggregates of 4 with
            // omega 2.5 / 1.5: 152 / 94 its, 93 / 59 ms; smoothed (0.9) aggregates of 4, omega 1.25: 40 its, 59 ms (denser
            // coarse operators eat the gain); smoothed aggregates of 8, omega 1.25: 59 its, 41.5 ms.
            if (sys.cfg.amg_smooth < 0.0) sys.cfg.amg_smooth = 0.9;
            if (sys.cfg.amg_passes <= 0) sys.cfg.amg_passes = 3;
            sys.own_P = aggregation_chain(S1, 64, 16, sys.cfg.amg_smooth, sys.cfg.amg_passes);
            Ps.clear();
            for (const HCsr &P : sys.own_P) Ps.push_back(&P);
            if (!sys.cfg.omega_user) sys.cfg.omega = sys.cfg.amg_smooth > 0.0 ? 1.25 : 1.5;
        } else if (sys.cfg.p_smooth > 0.0 && !Ps.empty()) {
            // the hierarchy's own agglomerates, prolongators smoothed like the aggregation path's (fixed across realisations)
            HCsr Sl = S1;
            for (const HCsr *P0 : std::vector<const HCsr *>(Ps)) {
                if (P0->rows != Sl.rows) break;
                HCsr P = smooth_prolongator(Sl, *P0, sys.cfg.p_smooth, 0.25);
                Sl = csr_matmul(csr_transpose(P), csr_matmul(Sl, P));
                sys.own_P.push_back(std::move(P));
            }
            if (sys.own_P.size() == Ps.size()) {
                Ps.clear();
                for (const HCsr &P : sys.own_P) Ps.push_back(&P);
            } else sys.own_P.clear();
        }
    }
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
            if ((rc = upload_wcsr(c, Sw, V.S, -1))) return rc;
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
    if (!sys.cfg.coarse_user && sys.own_P.empty()) { sys.cfg.coarse_degree = 6; sys.cfg.coarse_ratio = 20.0; }
    {
        std::vector<double> cc = cheb_coefficients(1.0 / sys.cfg.coarse_ratio, 1.0, sys.cfg.coarse_degree);
        if ((rc = to_device(c, cc, &sys.d_coarse_coef))) return rc;
    }
    if (sys.cfg.schur_degree_coarse < 0) sys.cfg.schur_degree_coarse = sys.own_P.empty() ? 3 : 0;
    sys.ready = true;
    return PMC_OK;
}

// --------------------------------------------------------------------------------------------------
// program builder: the host records the operations of a level batch once; k_run_program executes them per tile
// --------------------------------------------------------------------------------------------------
typedef long long Off;  // offset (in doubles) inside a tile chunk; < 0: none
static VecRef vr(Off off, int /*nrows*/ = 0, int row_off = 0)
{
    VecRef v;
    v.off = off >= 0 ? off + (Off)row_off * TW : -1;
    return v;
}
static const VecRef VNULL = {-1};

// Row allocator of the tile chunk: every tile owns `peak` doubles laid out identically.
struct Rows {
    Off top = 0, peak = 0;
    Off alloc(long long nrows)
    {
        const Off o = top;
        top += nrows * TW;
        if (top > peak) peak = top;
        return o;
    }
};

struct Program {
    std::vector<Op> ops;
    bool staging = true;  // stage operator entries through shared memory where the slices fit (F_STAGED)
    bool defer_x = true;  // MINRES: apply the solution updates of an iteration pair in one pass (option "defer_x")
    bool cheb3 = true;    // sampler Chebyshev semi-iteration in three-term form (option "cheb_three_term")
    bool qoi_only = true; // Darcy MINRES carries obs . x by scalar recurrences when only Q is returned (option "qoi_only")
    bool fuse_coarse = true;  // coarsest Chebyshev iteration as one shared-memory operation (option "fuse_coarse")
    bool split_apply = true;  // Darcy block operator applied as two operations (option "split_apply")
    bool chunked = true;      // wide slices staged chunk by chunk instead of read from L2 (option "stage_wide")
    int pc() const { return (int)ops.size(); }
    Op &add(int kind, int kclass, int n, double rows_moved, double matrix_bytes = 0.0)
    {
        Op o;
        memset(&o, 0, sizeof o);
        o.kind = kind;
        o.kclass = kclass;
        o.n = n;
        o.bytes = rows_moved * TW * 8.0;  // batched-vector traffic of one tile; the shared matrices stay in L2
        (void)matrix_bytes;
        o.x = o.y = o.r = o.d = o.w = o.v = VNULL;
        ops.push_back(o);
        return ops.back();
    }
};

static void emit_spmm(Program &pg, int kclass, int ep, const DevCsr &A, VecRef V, VecRef x, VecRef y, VecRef r, VecRef d,
                      const double *dinv_fixed, VecRef dinv_b, double ca, double cb, int dot_slot, bool dot_acc,
                      bool dot_with_r, double rows_moved)
{
    Op &o = pg.add(OP_SPMM, kclass, A.rows, rows_moved, A.matrix_bytes());
    o.flags = (ep << F_EP_SHIFT) | (A.weighted ? F_WEIGHTED : 0) | ((ep == EP_CHEB && A.weighted) ? F_BDINV : 0) |
              (dot_slot >= 0 ? F_DOT : 0) | (dot_acc ? F_DOT_ACC : 0) | (dot_with_r ? F_DOT_WITH_R : 0) |
              ((pg.staging && A.max_width <= STW) ? F_STAGED : 0) |
              ((pg.staging && pg.chunked && A.max_width > STW) ? F_CHUNKED : 0);
    o.rowptr = A.soff; o.pk = A.spk;
    o.fixed = dinv_fixed;
    o.x = x; o.y = y; o.r = r; o.d = d; o.w = dinv_b; o.v = V;
    o.ca = ca; o.cb = cb;
    o.slot = dot_slot < 0 ? 0 : dot_slot;
}

static void emit_fill(Program &pg, VecRef y, int n, double v)
{
    Op &o = pg.add(OP_FILL, KC_MISC, n, n);
    o.y = y;
    o.ca = v;
}
static void emit_copy(Program &pg, VecRef x, VecRef y, int n)
{
    Op &o = pg.add(OP_COPY, KC_MISC, n, 2.0 * n);
    o.x = x;
    o.y = y;
}

struct ChebOp {
    const DevCsr *A;
    VecRef V;               // weights (weighted operators)
    const double *dinv_f;   // fixed [n]   (plain operators)
    VecRef dinv_b;          // batched     (weighted operators)
    double lo, hi;
    double vrows;           // weight rows read per apply
    int kclass;
};

// `deg` Chebyshev steps for A z = r.  from_zero: z_0 = 0 and the result ends in `cur`.  Otherwise the current iterate
// lives in `cur` and the result ends in (deg even ? cur : other).  Returns the buffer holding the result.
static VecRef emit_cheb(Program &pg, const ChebOp &op, VecRef r, VecRef d, int deg, bool from_zero, VecRef cur, VecRef other,
                        int dot_slot, bool dot_acc)
{
    const int n = op.A->rows;
    const bool bd = op.A->weighted;
    const double theta = 0.5 * (op.hi + op.lo), delta = 0.5 * (op.hi - op.lo), sigma = theta / delta;
    double rho = 1.0 / sigma;
    VecRef zin = cur, zout = other;
    int j0 = 0;
    if (from_zero) {
        zout = (deg % 2 == 1) ? cur : other;
        Op &o = pg.add(OP_CHEB_FIRST, op.kclass, n, n * (3.0 + (bd ? 1 : 0)));
        o.flags = (bd ? F_BDINV : 0) | ((dot_slot >= 0 && deg == 1) ? F_DOT : 0) | (dot_acc ? F_DOT_ACC : 0);
        o.r = r; o.d = d; o.y = zout; o.w = op.dinv_b; o.fixed = op.dinv_f;
        o.cb = 1.0 / theta;
        o.slot = dot_slot < 0 ? 0 : dot_slot;
        zin = zout;
        zout = (zin.off == cur.off) ? other : cur;
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
        const bool dl = dot_slot >= 0 && j == deg - 1;
        emit_spmm(pg, op.kclass, EP_CHEB, *op.A, op.V, zin, zout, r, d, op.dinv_f, op.dinv_b, ca, cb, dl ? dot_slot : -1,
                  dl && dot_acc, true, n * (5.0 + (bd ? 1 : 0)) + op.vrows);
        VecRef t = zin; zin = zout; zout = t;
    }
    return zin;
}

// --------------------------------------------------------------------------------------------------
// solve workspace (tile-major batched vectors: n rows -> n * ld doubles, ld = ntiles * TW)
// --------------------------------------------------------------------------------------------------
// {ca_j, cb_j}, j = 0..deg-1, of emit_cheb's recurrence (from_zero) on [lo, hi]
static std::vector<double> cheb_coefficients(double lo, double hi, int deg)
{
    const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
    double rho = 1.0 / sigma;
    std::vector<double> c(2 * (size_t)deg);
    for (int j = 0; j < deg; ++j) {
        if (j == 0) { c[0] = 0.0; c[1] = 1.0 / theta; }
        else {
            const double rho_new = 1.0 / (2.0 * sigma - rho);
            c[2 * j] = rho_new * rho;
            c[2 * j + 1] = 2.0 * rho_new / delta;
            rho = rho_new;
        }
    }
    return c;
}

struct SolveWs {
    Off v0, v1, w0, w1, u1, q, x, b;  // MINRES, N rows each
    Off mu_d, mu_z;                   // mass-block Chebyshev scratch, Nf rows
    Off dinvM;                        // Darcy: batched 1/diag M(k), Nf rows
    std::vector<Off> vr_, vzA, vzB, vd, vres, vV, vl1;  // per V-level
    Off iters;                        // one row: iteration counts of the tile's samples (as doubles)
};

static void carve_solve(Rows &ar, const SaddleSys &sys, SolveWs &ws)
{
    const long long N = sys.N, Nf = sys.Nf;
    ws.v0 = ar.alloc(N); ws.v1 = ar.alloc(N); ws.w0 = ar.alloc(N); ws.w1 = ar.alloc(N);
    ws.u1 = ar.alloc(N); ws.q = ar.alloc(N); ws.x = ar.alloc(N); ws.b = ar.alloc(N);
    ws.mu_d = ar.alloc(Nf);
    ws.mu_z = ar.alloc(Nf);
    ws.dinvM = sys.weighted ? ar.alloc(Nf) : -1;
    const size_t nv = sys.v.size();
    ws.vr_.assign(nv, -1); ws.vzA.assign(nv, -1); ws.vzB.assign(nv, -1); ws.vd.assign(nv, -1);
    ws.vres.assign(nv, -1); ws.vV.assign(nv, -1); ws.vl1.assign(nv, -1);
    for (size_t m = 0; m < nv; ++m) {
        const long long n = sys.v[m].n;
        if (m > 0) { ws.vr_[m] = ar.alloc(n); ws.vzA[m] = ar.alloc(n); }
        ws.vzB[m] = ar.alloc(n);
        ws.vd[m] = ar.alloc(n);
        if (m + 1 < nv) ws.vres[m] = ar.alloc(n);
        if (sys.weighted) { ws.vV[m] = ar.alloc(sys.v[m].nU); ws.vl1[m] = ar.alloc(n); }
    }
    ws.iters = ar.alloc(1);
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

struct Solver {
    SaddleSys *sys;
    SolveWs *ws;
    VecRef k_ext;  // Darcy weights [Ne+1 rows] or null
};

static void emit_vcycle(Program &pg, Solver &sv, int m, VecRef r, VecRef zout, VecRef ztmp, int dot_slot)
{
    SaddleSys &sys = *sv.sys;
    SolveWs &ws = *sv.ws;
    VLevel &L = sys.v[m];
    const PrecCfg &cfg = sys.cfg;
    ChebOp op;
    op.A = &L.S;
    op.V = sys.weighted ? vr(ws.vV[m], L.nU) : VNULL;
    op.dinv_f = sys.weighted ? nullptr : L.l1inv_fixed;
    op.dinv_b = sys.weighted ? vr(ws.vl1[m], L.n) : VNULL;
    op.hi = 1.0;
    op.vrows = sys.weighted ? L.nU : 0;
    op.kclass = KC_SCHUR;
    VecRef d = vr(ws.vd[m], L.n);
    const bool last = (m + 1 == (int)sys.v.size());
    if (last) {
        op.lo = 1.0 / cfg.coarse_ratio;
        if (pg.fuse_coarse && sys.d_coarse_coef && L.n <= SMALLN && L.S.max_width > 0) {
            // one operation for the whole iteration; the result always lands in zout
            // the iterates stay in shared memory: only r, 1/l1 and the weights are read and the result written, once
            Op &o = pg.add(OP_CHEB_SMALL, KC_SCHUR, L.n, L.n * (2.0 + (sys.weighted ? 1 : 0)) + op.vrows);
            o.flags = (sys.weighted ? F_WEIGHTED : 0) | (dot_slot >= 0 ? (F_DOT | F_DOT_ACC) : 0);
            o.rowptr = L.S.soff; o.pk = L.S.spk; o.val = sys.d_coarse_coef; o.fixed = op.dinv_f;
            o.r = r; o.y = zout; o.v = op.V; o.w = op.dinv_b;
            o.a0 = cfg.coarse_degree;
            o.a1 = sys.weighted ? L.nU : 0;   // weight rows (staged in shared memory with the operator's entries when they fit)
            o.slot = dot_slot < 0 ? 0 : dot_slot;
            return;
        }
        emit_cheb(pg, op, r, d, cfg.coarse_degree, true, zout, ztmp, dot_slot, dot_slot >= 0);
        return;
    }
    op.lo = 1.0 / cfg.schur_ratio;
    const int s = (m > 0 && cfg.schur_degree_coarse > 0) ? cfg.schur_degree_coarse : cfg.schur_degree;
    VecRef E = (s % 2 == 0) ? zout : ztmp, O = (s % 2 == 0) ? ztmp : zout;
    emit_cheb(pg, op, r, d, s, true, E, O, -1, false);
    VLevel &Lc = sys.v[m + 1];
    VecRef res = vr(ws.vres[m], L.n), rc = vr(ws.vr_[m + 1], Lc.n), zc = vr(ws.vzA[m + 1], Lc.n), zct = vr(ws.vzB[m + 1], Lc.n);
    emit_spmm(pg, KC_SCHUR, EP_RESID, L.S, op.V, E, res, r, VNULL, nullptr, VNULL, 0, 0, -1, false, false, 3.0 * L.n + op.vrows);
    emit_spmm(pg, KC_TRANSFER, EP_AX, L.Pt, VNULL, res, rc, VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false, (double)L.n + Lc.n);
    emit_vcycle(pg, sv, m + 1, rc, zc, zct, -1);
    emit_spmm(pg, KC_TRANSFER, EP_ADD, L.P, VNULL, zc, E, VNULL, VNULL, nullptr, VNULL, cfg.omega, 0, -1, false, false,
              2.0 * L.n + Lc.n);
    emit_cheb(pg, op, r, d, s, false, E, O, dot_slot, dot_slot >= 0);
}

// z = Prec r (block diagonal: Chebyshev-Jacobi on the RT mass block, V-cycle on the Schur complement); if dot_slot >= 0
// the slot receives r . z (mass part assigns, Schur part accumulates).
static void emit_prec(Program &pg, Solver &sv, Off r, Off z, int dot_slot, bool mass_done = false)
{
    SaddleSys &sys = *sv.sys;
    SolveWs &ws = *sv.ws;
    if (mass_done) {  // the mass-block Jacobi (and its share of the dot) was fused into the Lanczos update
        emit_vcycle(pg, sv, 0, vr(r, sys.N, sys.Nf), vr(z, sys.N, sys.Nf), vr(ws.vzB[0], sys.Ne), dot_slot);
        return;
    }
    ChebOp op;
    op.A = &sys.Muu;
    op.V = sv.k_ext;
    op.dinv_f = sys.weighted ? nullptr : sys.dinvM_fixed;
    op.dinv_b = sys.weighted ? vr(ws.dinvM, sys.Nf) : VNULL;
    const double ms = sys.cfg.mass_degree == 1 ? sys.cfg.mass_scale : 1.0;
    op.lo = sys.m_lo * ms;
    op.hi = sys.m_hi * ms;
    op.vrows = sys.weighted ? sys.Ne : 0;
    op.kclass = KC_MASS;
    emit_cheb(pg, op, vr(r, sys.N), vr(ws.mu_d, sys.Nf), sys.cfg.mass_degree, true, vr(z, sys.N), vr(ws.mu_z, sys.Nf), dot_slot,
              false);
    emit_vcycle(pg, sv, 0, vr(r, sys.N, sys.Nf), vr(z, sys.N, sys.Nf), vr(ws.vzB[0], sys.Ne), dot_slot);
}

static void emit_saddle(Program &pg, Solver &sv, int ep, Off x, Off y, Off r, int dot_slot)
{
    SaddleSys &sys = *sv.sys;
    if (ep == EP_AX && sys.split_apply && pg.split_apply) {
        // q_u = [M(k) | B^T] u over the RT rows (weighted), q_p = B u_u over the pressure rows (plain); the fused dot u . q
        // accumulates over the two operations
        emit_spmm(pg, KC_SADDLE, EP_AX, sys.Au, sv.k_ext, vr(x, sys.N), vr(y, sys.N), VNULL, VNULL, nullptr, VNULL, 0, 0, dot_slot,
                  false, false, (double)sys.N + sys.Nf + sys.Ne);
        emit_spmm(pg, KC_SADDLE, EP_AX, sys.Bp, VNULL, vr(x, sys.N), vr(y, sys.N, sys.Nf), dot_slot >= 0 ? vr(x, sys.N, sys.Nf) : VNULL,
                  VNULL, nullptr, VNULL, 0, 0, dot_slot, dot_slot >= 0, false, (double)sys.Ne);   // u was credited above
        return;
    }
    emit_spmm(pg, KC_SADDLE, ep, sys.A, sv.k_ext, vr(x, sys.N), vr(y, sys.N), r >= 0 ? vr(r, sys.N) : VNULL, VNULL, nullptr, VNULL, 0,
              0, dot_slot, false, false, (ep == EP_RESID ? 3.0 : 2.0) * sys.N + (sys.weighted ? sys.Ne : 0));
}

// Darcy per-solve values: 1/diag M(k), Schur values on every V-level, l1 norms.
static void emit_darcy_setup(Program &pg, Solver &sv)
{
    SaddleSys &sys = *sv.sys;
    SolveWs &ws = *sv.ws;
    {
        Op &o = pg.add(OP_SETUP_SPMM, KC_SETUP, sys.Nf, (double)sys.Ne + sys.Nf, sys.Dm.matrix_bytes());
        o.flags = F_RECIP;
        o.rowptr = sys.Dm.rowptr; o.col = sys.Dm.col; o.val = sys.Dm.val;
        o.x = sv.k_ext;
        o.y = vr(ws.dinvM, sys.Nf);
    }
    VecRef prev = vr(ws.dinvM, sys.Nf);
    int prev_rows = sys.Nf;
    for (size_t m = 0; m < sys.v.size(); ++m) {
        VLevel &L = sys.v[m];
        {
            Op &o = pg.add(OP_SETUP_SPMM, KC_SETUP, L.nU, (double)prev_rows + L.nU, L.T.matrix_bytes());
            o.rowptr = L.T.rowptr; o.col = L.T.col; o.val = L.T.val;
            o.x = prev;
            o.y = vr(ws.vV[m], L.nU);
        }
        {
            Op &o = pg.add(OP_SETUP_SPMM, KC_SETUP, L.n, (double)L.nU + L.n, L.L.matrix_bytes());
            o.flags = F_ABSX | F_RECIP;
            o.rowptr = L.L.rowptr; o.col = L.L.col; o.val = L.L.val;
            o.x = vr(ws.vV[m], L.nU);
            o.y = vr(ws.vl1[m], L.n);
        }
        prev = vr(ws.vV[m], L.nU);
        prev_rows = L.nU;
    }
}

// Preconditioned MINRES; ws.b and ws.x are set by earlier operations (x_nonzero: x holds an initial guess).
// qoi (with qoi_row): only Q = qoi . x is wanted and x starts from 0: the functional is carried by scalar recurrences (see
// sc_beta) driven by qoi . z_k, one sparse dot per iteration; the direction vectors w and the solution x are never formed
// (no OP_SOL_UPDATE: 5 N of the ~21 N rows an iteration moves).  qoi_rows = number of non-zero entries of qoi.
static void emit_minres(Program &pg, Solver &sv, bool x_nonzero, bool store_iters, const double *qoi = nullptr, Off qoi_row = -1,
                        int qoi_rows = 0, const int *qoi_idx = nullptr)
{
    SaddleSys &sys = *sv.sys;
    SolveWs &ws = *sv.ws;
    const int N = sys.N;
    const bool functional = qoi != nullptr && qoi_idx != nullptr && !x_nonzero;
    if (x_nonzero) emit_saddle(pg, sv, EP_RESID, ws.x, ws.v1, ws.b, -1);
    else emit_copy(pg, vr(ws.b, N), vr(ws.v1, N), N);
    emit_fill(pg, vr(ws.v0, N), N, 0.0);
    if (!functional) {
        emit_fill(pg, vr(ws.w0, N), N, 0.0);
        emit_fill(pg, vr(ws.w1, N), N, 0.0);
    }
    emit_prec(pg, sv, ws.v1, ws.u1, 1);
    { Op &o = pg.add(OP_SC_INIT, KC_SCALAR, 0, 0); o.slot = 1; }
    std::vector<int> exits_a, exits_b;
    const bool defer_x = pg.defer_x && !functional;
    exits_b.push_back(pg.pc());
    pg.add(OP_CHECK, KC_SCALAR, 0, 0);
    const int loop_start = pg.pc();
    Off v0 = ws.v0, v1 = ws.v1, w0 = ws.w0, w1 = ws.w1, u1 = ws.u1, q = ws.q;
    for (int parity = 0; parity < 2; ++parity) {
        emit_saddle(pg, sv, EP_AX, u1, q, -1, 0);
        if (functional) {   // dots[3] = qoi . z_k over the non-zero entries of the functional (qoi = their values, qoi_idx their rows)
            Op &o = pg.add(OP_DOT_SPARSE, KC_SOLUPD, qoi_rows, (double)qoi_rows);
            o.val = qoi;
            o.col = qoi_idx;
            o.x = vr(u1, N);
        }
        { Op &o = pg.add(OP_SC_ALPHA, KC_SCALAR, 0, 0); o.slot = 0; }
        const bool fuse_jacobi = sys.cfg.mass_degree == 1;
        {
            Op &o = pg.add(OP_LINCOMB3, KC_LANCZOS, N, 4.0 * N + (fuse_jacobi ? (sys.weighted ? 2.0 : 1.0) * sys.Nf : 0.0));
            o.x = vr(q, N); o.r = vr(v1, N); o.y = vr(v0, N);
            if (fuse_jacobi) {
                // z_u = dinvM v0_u / theta written straight into the u block of q (q's own rows are read first)
                const double theta = 0.5 * (sys.m_hi + sys.m_lo) * sys.cfg.mass_scale;
                o.flags = F_DOT | (sys.weighted ? F_BDINV : 0);
                o.d = vr(q, N);
                o.w = sys.weighted ? vr(ws.dinvM, sys.Nf) : VNULL;
                o.fixed = sys.weighted ? nullptr : sys.dinvM_fixed;
                o.cb = 1.0 / theta;
                o.a0 = sys.Nf;
                o.slot = 1;
            }
        }
        emit_prec(pg, sv, v0, q, 1, fuse_jacobi);
        // The solution update x += cx w of the first iteration of a pair is deferred and applied together with the
        // second one's (which reads that direction vector anyway): one read and one write of x less per pair, the same
        // floating-point operations in the same order.
        { Op &o = pg.add(OP_SC_BETA, KC_SCALAR, 0, 0); o.slot = 1; o.a0 = defer_x && parity == 0 ? 1 : 0; o.a1 = functional ? 1 : 0; }
        if (!functional) {
            const int mode = !defer_x ? 0 : (parity == 0 ? 1 : 2);
            Op &o = pg.add(OP_SOL_UPDATE, KC_SOLUPD, N, mode == 1 ? 4.0 * N : 6.0 * N);
            o.y = vr(w0, N); o.r = vr(w1, N); o.x = vr(u1, N); o.d = vr(ws.x, N);
            o.a0 = mode;
        }
        (parity == 0 ? exits_a : exits_b).push_back(pg.pc());
        pg.add(OP_CHECK, KC_SCALAR, 0, 0);
        std::swap(u1, q);
        std::swap(v0, v1);
        std::swap(w0, w1);
    }
    { Op &o = pg.add(OP_JUMP, KC_SCALAR, 0, 0); o.a0 = loop_start; }
    for (int i = loop_start; i < pg.pc(); ++i) pg.ops[i].flags |= F_INLOOP;
    // leaving after the first iteration of a pair: apply its deferred update (w0 of that iteration = ws.w0)
    const int exit_a = pg.pc();
    if (defer_x) { Op &o = pg.add(OP_SOL_UPDATE, KC_SOLUPD, N, 3.0 * N); o.y = vr(ws.w0, N); o.d = vr(ws.x, N); o.a0 = 3; }
    const int exit_b = pg.pc();
    for (int e : exits_a) pg.ops[e].a0 = exit_a;
    for (int e : exits_b) pg.ops[e].a0 = exit_b;
    if (functional) { Op &o = pg.add(OP_STORE_Q, KC_SCALAR, 0, 0); o.y = vr(qoi_row); }
    { Op &o = pg.add(OP_STORE_ITERS, KC_SCALAR, 0, 0); o.y = store_iters ? vr(ws.iters) : VNULL; }
}

// Sampler solve: rhs_p = batched right-hand side at `level` (Ne rows); x0_p (nullable) initial guess of the Gaussian
// field.  The field is left in rows [Nf, N) of ws.x.
// Jacobi-preconditioned CG on the SPD form (see prepare_sampler).  Workspace: [p; t] = ws.u1 (N rows: the direction and
// t = Bs p), q = H p in ws.q, the residual in ws.v0, the solution u in rows [0, Nf) of ws.x; the field
// s = Bs u - f / (alpha W) is left in rows [Nf, N) of ws.x like the MINRES path leaves it.  Stopping rule: the
// preconditioned residual norm sqrt(r . D^-1 r) <= max(rel * its initial value, abs), per realisation.
static void emit_sampler_pcg(Program &pg, SaddleSys &sys, Off rhs_p, SolveWs &ws, bool store_iters)
{
    const int Nf = sys.Nf, Ne = sys.Ne, N = sys.N;
    const VecRef P = vr(ws.u1, N), Pt = vr(ws.u1, N, Nf), Q = vr(ws.q, N), R = vr(ws.v0, N), X = vr(ws.x, N), Sx = vr(ws.x, N, Nf);
    auto scale = [&](VecRef dst, double cb) {   // dst = cb * f / (alpha W)
        Op &o = pg.add(OP_CHEB_FIRST, KC_MISC, Ne, 2.0 * Ne);
        o.r = vr(rhs_p, Ne); o.d = dst; o.y = dst; o.fixed = sys.inv_aw; o.cb = cb;
    };
    emit_fill(pg, P, Nf, 0.0);
    scale(Pt, 1.0);
    scale(Sx, -1.0);
    // r0 = alpha^-1 B^T W^-1 f = [M | B^T] [0; f / (alpha W)]
    emit_spmm(pg, KC_SADDLE, EP_AX, sys.MBt, VNULL, P, R, VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false, (double)Ne + Nf);
    {   // p0 = z0 = D^-1 r0, dots[1] = r0 . z0
        Op &o = pg.add(OP_CHEB_FIRST, KC_MASS, Nf, 2.0 * Nf);
        o.flags = F_DOT;
        o.r = R; o.d = P; o.y = P; o.fixed = sys.dinvH; o.cb = 1.0; o.slot = 1;
    }
    emit_fill(pg, X, Nf, 0.0);
    { Op &o = pg.add(OP_CG_INIT, KC_SCALAR, 0, 0); o.slot = 1; }
    const int loop_start = pg.pc();
    const int check = pg.pc();
    pg.add(OP_CHECK, KC_SCALAR, 0, 0);
    emit_spmm(pg, KC_SADDLE, EP_AX, sys.Bs, VNULL, P, Pt, VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false, (double)Nf + Ne);
    emit_spmm(pg, KC_SADDLE, EP_AX, sys.MBt, VNULL, P, Q, VNULL, VNULL, nullptr, VNULL, 0, 0, 0, false, false, (double)N + Nf);
    { Op &o = pg.add(OP_CG_ALPHA, KC_SCALAR, 0, 0); o.slot = 0; }
    {
        Op &o = pg.add(OP_CG_UPDATE, KC_SOLUPD, Nf, 6.0 * Nf);
        o.x = P; o.r = Q; o.y = R; o.d = X; o.fixed = sys.dinvH; o.slot = 1;
    }
    { Op &o = pg.add(OP_CG_BETA, KC_SCALAR, 0, 0); o.slot = 1; }
    {
        Op &o = pg.add(OP_CG_DIR, KC_LANCZOS, Nf, 3.0 * Nf);
        o.x = R; o.y = P; o.fixed = sys.dinvH;
    }
    { Op &o = pg.add(OP_JUMP, KC_SCALAR, 0, 0); o.a0 = loop_start; }
    for (int i = loop_start; i < pg.pc(); ++i) pg.ops[i].flags |= F_INLOOP;
    pg.ops[check].a0 = pg.pc();
    // s = Bs u - f / (alpha W)
    emit_spmm(pg, KC_SADDLE, EP_ADD, sys.Bs, VNULL, X, Sx, VNULL, VNULL, nullptr, VNULL, 1.0, 0, -1, false, false, (double)Nf + 2.0 * Ne);
    { Op &o = pg.add(OP_STORE_ITERS, KC_SCALAR, 0, 0); o.y = store_iters ? vr(ws.iters) : VNULL; }
}

// Chebyshev semi-iteration on the SPD form H u = b0 (same system and same stopping rule as emit_sampler_pcg).  H does not
// depend on the realisation, so the spectrum [lo, hi] of diag(H)^-1 H is computed once at set-up and m Chebyshev steps
// reduce the D^-1-norm of every residual by at least 1 / T_m((hi + lo) / (hi - lo)) <= rel: the step count is known
// before the solve, no step needs a dot product, and a step is two sparse applies with the whole update in the second
// one's epilogue (t = Bs z; d = ca d + cb D^-1 (b0 - [M | B^T][z; t]); z' = z + d): N + 4 Nf + (Nf + Ne) rows per step
// against 12 Nf + 2 Ne for a PCG iteration.  After the m steps the true residual norm is evaluated and checked per
// realisation; tiles that still hold an unconverged realisation run further (restarted) blocks of steps.
// Workspace: the iterates ping-pong between [z; t] = ws.u1 and ws.q (N rows each), d = ws.w0, b0 = ws.v0, the residual
// of the check = ws.w1; the field s = Bs u - f / (alpha W) is left in rows [Nf, N) of ws.x like the other paths leave it.
static void emit_sampler_cheb(Program &pg, SaddleSys &sys, Off rhs_p, SolveWs &ws, bool store_iters, double rel, int maxit)
{
    const int Nf = sys.Nf, Ne = sys.Ne, N = sys.N;
    VecRef Z[2] = {vr(ws.u1, N), vr(ws.q, N)}, Zt[2] = {vr(ws.u1, N, Nf), vr(ws.q, N, Nf)};
    const VecRef R = vr(ws.v0, N), D = vr(ws.w0, N), RES = vr(ws.w1, N), Sx = vr(ws.x, N, Nf);
    const double lo = sys.h_lo, hi = sys.h_hi;
    const double theta = 0.5 * (hi + lo), sigma = (hi + lo) / (hi - lo);
    const double want = rel > 0.0 && rel < 1.0 ? rel : 1e-16;
    int m = (int)std::ceil(std::acosh(1.0 / want) / std::acosh(sigma));
    m = std::max(2, std::min(m, std::max(2, maxit)));
    auto scale = [&](VecRef dst, double cb) {   // dst = cb * f / (alpha W)
        Op &o = pg.add(OP_CHEB_FIRST, KC_MISC, Ne, 2.0 * Ne);
        o.r = vr(rhs_p, Ne); o.d = dst; o.y = dst; o.fixed = sys.inv_aw; o.cb = cb;
    };
    // Three-term form (default): the update d = z - z_prev is not stored; the step reads the previous iterate from the
    // buffer it overwrites: z' = z + ca (z - z_prev) + cb D^-1 (b0 - H z), one row pass less per step.
    const bool three = pg.cheb3;
    auto step = [&](int cur, double ca, double cb) {   // Z[1 - cur] = Z[cur] + d,  d = ca d + cb D^-1 (b0 - H Z[cur])
        emit_spmm(pg, KC_SADDLE, EP_AX, sys.Bs, VNULL, Z[cur], Zt[cur], VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false, (double)Nf + Ne);
        emit_spmm(pg, KC_SADDLE, EP_CHEB, sys.MBt, VNULL, Z[cur], Z[1 - cur], R, three ? VNULL : D, sys.dinvH, VNULL, ca, cb, -1, false,
                  false, (double)N + (three ? 3.0 : 4.0) * Nf - (ca == 0.0 ? Nf : 0));
        if (three) pg.ops.back().flags |= F_THREE;
    };
    auto check = [&](int cur, int steps_done) {   // dots[1] = r . D^-1 r with r = b0 - H Z[cur]; convergence per realisation
        emit_spmm(pg, KC_SADDLE, EP_AX, sys.Bs, VNULL, Z[cur], Zt[cur], VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false, (double)Nf + Ne);
        emit_spmm(pg, KC_SADDLE, EP_RESID, sys.MBt, VNULL, Z[cur], RES, R, VNULL, nullptr, VNULL, 0, 0, -1, false, false, (double)N + 2.0 * Nf);
        {
            Op &o = pg.add(OP_CHEB_FIRST, KC_MASS, Nf, 2.0 * Nf);   // the scaled residual lands in the idle iterate buffer
            o.flags = F_DOT;
            o.r = RES; o.d = Z[1 - cur]; o.y = Z[1 - cur]; o.fixed = sys.dinvH; o.cb = 1.0; o.slot = 1;
        }
        { Op &o = pg.add(OP_CHB_CHECK, KC_SCALAR, 0, 0); o.slot = 1; o.a0 = steps_done; }
    };
    emit_fill(pg, Z[0], Nf, 0.0);
    if (three) emit_fill(pg, Z[1], Nf, 0.0);   // z_0 = 0 is the "previous iterate" of the second step
    scale(Zt[0], 1.0);
    scale(Sx, -1.0);
    // b0 = alpha^-1 B^T W^-1 f = [M | B^T] [0; f / (alpha W)]
    emit_spmm(pg, KC_SADDLE, EP_AX, sys.MBt, VNULL, Z[0], R, VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false, (double)Ne + Nf);
    {   // first step from u = 0: z1 = d = D^-1 b0 / theta;  dots[1] = b0 . D^-1 b0 / theta
        Op &o = pg.add(OP_CHEB_FIRST, KC_MASS, Nf, (three ? 2.0 : 3.0) * Nf);
        o.flags = F_DOT;
        o.r = R; o.d = three ? Z[0] : D; o.y = Z[0]; o.fixed = sys.dinvH; o.cb = 1.0 / theta; o.slot = 1;
    }
    { Op &o = pg.add(OP_CHB_INIT, KC_SCALAR, 0, 0); o.slot = 1; o.ca = theta; }
    const int loop_start = pg.pc();
    std::vector<int> exits;
    exits.push_back(pg.pc());
    pg.add(OP_CHECK, KC_SCALAR, 0, 0);
    std::vector<double> cf = cheb_coefficients(lo, hi, m);
    int cur = 0;
    for (int j = 1; j < m; ++j) { step(cur, cf[2 * j], cf[2 * j + 1]); cur ^= 1; }
    check(cur, m);
    exits.push_back(pg.pc());
    pg.add(OP_CHECK, KC_SCALAR, 0, 0);
    // restarted blocks (an even number of steps, so that the iterate returns to the same buffer) until every realisation of
    // the tile meets the stopping rule; only reached if the spectrum estimate was too narrow
    const int m2 = std::max(2, 2 * ((m + 3) / 4));
    std::vector<double> cf2 = cheb_coefficients(lo, hi, m2);
    const int restart = pg.pc();
    for (int j = 0; j < m2; ++j) { step(cur, cf2[2 * j], cf2[2 * j + 1]); cur ^= 1; }
    check(cur, m2);
    exits.push_back(pg.pc());
    pg.add(OP_CHECK, KC_SCALAR, 0, 0);
    { Op &o = pg.add(OP_JUMP, KC_SCALAR, 0, 0); o.a0 = restart; }
    for (int i = loop_start; i < pg.pc(); ++i) pg.ops[i].flags |= F_INLOOP;
    for (int e : exits) pg.ops[e].a0 = pg.pc();
    // s = Bs u - f / (alpha W)
    emit_spmm(pg, KC_SADDLE, EP_ADD, sys.Bs, VNULL, Z[cur], Sx, VNULL, VNULL, nullptr, VNULL, 1.0, 0, -1, false, false, (double)Nf + 2.0 * Ne);
    { Op &o = pg.add(OP_STORE_ITERS, KC_SCALAR, 0, 0); o.y = store_iters ? vr(ws.iters) : VNULL; }
}

static void emit_sampler_solve(Program &pg, Ctx *c, int level, Off rhs_p, Off x0_p, SolveWs &ws, bool store_iters)
{
    SaddleSys &sys = c->s[level].sys;
    if (sys.pcg && sys.cheb) {
        emit_sampler_cheb(pg, sys, rhs_p, ws, store_iters, c->rel, c->maxit);
        return;
    }
    if (sys.pcg) {   // starts from u = 0: the prolongated coarse field (x0_p) is an initial guess for s only
        emit_sampler_pcg(pg, sys, rhs_p, ws, store_iters);
        return;
    }
    Solver sv{&sys, &ws, VNULL};
    emit_fill(pg, vr(ws.b, sys.N), sys.Nf, 0.0);  // rhs_u = 0 (/root/reference/src/PDESampler.cpp:441-442)
    emit_copy(pg, vr(rhs_p, sys.Ne), vr(ws.b, sys.N, sys.Nf), sys.Ne);
    emit_fill(pg, vr(ws.x, sys.N), sys.Nf, 0.0);
    if (x0_p >= 0) emit_copy(pg, vr(x0_p, sys.Ne), vr(ws.x, sys.N, sys.Nf), sys.Ne);
    else emit_fill(pg, vr(ws.x, sys.N, sys.Nf), sys.Ne, 0.0);
    emit_minres(pg, sv, x0_p >= 0, store_iters);
}

// Darcy solve: k_ext = batched [Ne+1 rows] (row Ne = 1).  Q_dev[sample] receives obs . sol.
// need_sol = false: the caller reads Q only (SolveFwd), so the solve may skip forming the solution (option "qoi_only").
static void emit_darcy_solve(Program &pg, Ctx *c, int level, Off k_ext, SolveWs &ws, Off Q_row, bool store_iters, bool need_sol = true)
{
    DarcyLevel &L = c->d[level];
    SaddleSys &sys = L.sys;
    Solver sv{&sys, &ws, vr(k_ext, sys.Ne + 1)};
    const int N = sys.N;
    emit_darcy_setup(pg, sv);
    { Op &o = pg.add(OP_BROADCAST, KC_MISC, N, N); o.fixed = L.d_rhs_bc; o.y = vr(ws.b, N); }
    if (L.ess_nonzero) {
        // rhs_bc -= M(k)[:, ess] ess_data  (BlockMatrix::EliminateRowCol, /root/reference/src/DarcySolver.cpp:498)
        { Op &o = pg.add(OP_BROADCAST, KC_MISC, sys.Nf, sys.Nf); o.fixed = L.d_ess_u_data; o.y = vr(ws.mu_z, sys.Nf); }
        emit_spmm(pg, KC_SADDLE, EP_RESID, L.Mbc, sv.k_ext, vr(ws.mu_z, sys.Nf), vr(ws.b, N), vr(ws.b, N), VNULL, nullptr, VNULL, 0,
                  0, -1, false, false, 3.0 * sys.Nf + sys.Ne);
    }
    if (!need_sol && Q_row >= 0 && pg.qoi_only) {
        emit_minres(pg, sv, false, store_iters, L.d_obs_val, Q_row, L.obs_nnz, L.d_obs_idx);   // p_sol = 0 (:629), Q = obs . sol (:427)
        return;
    }
    emit_fill(pg, vr(ws.x, N), N, 0.0);  // p_sol = 0 (:629)
    emit_minres(pg, sv, false, store_iters);
    if (Q_row >= 0) {
        Op &o = pg.add(OP_DOT_FIXED, KC_MISC, N, N);  // Q = obs . sol (:427)
        o.fixed = L.d_obs;
        o.x = vr(ws.x, N);
        o.y = vr(Q_row);
    }
}

static int pad_ld(int nsamples) { return ((nsamples + TW - 1) / TW) * TW; }

static int pick_batch(Ctx *c, size_t bytes_per_sample, int nsamples)
{
    int b = c->max_batch;
    if (b <= 0 && c->arena.cap >= bytes_per_sample * (size_t)pad_ld(nsamples) + (size_t)nsamples * 32 + (1 << 16)) b = nsamples;
    if (b <= 0) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        free_b += c->arena.cap;
        const size_t usable = (size_t)(0.80 * (double)free_b);
        size_t nb = usable / std::max<size_t>(bytes_per_sample, 1);
        if (nb > 16384) nb = 16384;
        b = (int)nb;
    }
    if (b >= TW) b = (b / TW) * TW;
    if (b < 1) b = 1;
    return std::min(b, std::max(nsamples, 1));
}

// Size the batch and its workspace in one step.  The managers run one handle per level from concurrent host threads on
// the same device: the free-memory reading and the allocation it leads to are made under one process-wide lock, so two
// handles never size their batches against the same free bytes, and an allocation that still fails (another process on
// the device) is retried with half the batch.  `per_sample` bytes per realisation, `extra` bytes per realisation of
// per-sample results, both for padded batches.
static std::mutex g_arena_mutex;
static int size_batch(Ctx *c, size_t per_sample, size_t extra, int nsamples, int *B_out)
{
    std::lock_guard<std::mutex> lk(g_arena_mutex);
    int B = pick_batch(c, per_sample, nsamples);
    for (;;) {
        int rc = ensure_arena(c, per_sample * (size_t)pad_ld(B) + (size_t)B * extra + (1 << 16));
        if (rc == PMC_OK) { *B_out = B; return PMC_OK; }
        if (rc != PMC_ERR_NOMEM || B <= TW) return rc;
        B = std::max(TW, ((B / 2) / TW) * TW);
    }
}

static int check_level(Ctx *c, int level, bool sampler, bool darcy)
{
    if (!c) return PMC_ERR_ARG;
    CK(cudaSetDevice(c->device));  // the current device is per host thread; callers may be pool threads
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

// Upload the program and run it: one CTA per tile, one launch.
template <int NTt, int MINB, int CS>
static cudaError_t launch_program(const ProgParams &P, int ntiles, cudaStream_t stream)
{
    // dynamic shared memory: the per-warp operator staging buffers (program.cuh)
    const size_t dyn = (size_t)(NTt / 32) * NSTAGE * sizeof(WarpStage);
    // the attribute is per (instantiation, device); level batches are launched from concurrent host threads, so the
    // "already set" cache is an atomic bit mask over devices (setting the attribute twice is harmless)
    static std::atomic<unsigned long long> attr_mask{0ull};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_mask.load(std::memory_order_acquire) & bit)) {
        cudaError_t e = cudaFuncSetAttribute(k_run_program<NTt, MINB, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        attr_mask.fetch_or(bit, std::memory_order_release);
    }
    if constexpr (CS == 1) {
        k_run_program<NTt, MINB, 1><<<ntiles, NTt, dyn, stream>>>(P);
        return cudaPeekAtLastError();
    } else if constexpr (CS == 0) {  // grid groups: cooperative launch, P.group CTAs per tile, all co-resident
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(ntiles * P.group));
        cfg.blockDim = dim3(NTt);
        cfg.dynamicSmemBytes = dyn;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k_run_program<NTt, MINB, 0>, P);
    } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ntiles * CS));
    cfg.blockDim = dim3(NTt);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_run_program<NTt, MINB, CS>, P);
    }
}

static int run_program(Ctx *c, Program &pg, int nsamples, Off chunk, int max_rows)
{
    const int ntiles = (nsamples + TW - 1) / TW;
    if (ntiles == 0 || pg.ops.empty()) return PMC_OK;
    const size_t bytes = pg.ops.size() * sizeof(Op);
    if (bytes > c->ops_cap) {
        if (c->d_ops) { cudaStreamSynchronize(c->stream); cudaFree(c->d_ops); cudaFreeHost(c->h_ops); }
        c->ops_cap = bytes * 2;
        CK(cudaMalloc((void **)&c->d_ops, c->ops_cap));
        CK(cudaMallocHost((void **)&c->h_ops, c->ops_cap));
    } else {
        CK(cudaStreamSynchronize(c->stream));  // the pinned staging buffer may still be in use by the previous copy
    }
    ProgParams P;
    P.ops = c->d_ops;
    P.nops = (int)pg.ops.size();
    P.ntiles = ntiles;
    P.nsamples = nsamples;
    P.max_iter = c->maxit;
    P.rel = c->rel;
    P.abs_ = c->abs_;
    P.mu = c->mu;
    P.sigma = c->sigma;
    P.tab = c->d_tab;
    P.stats = c->d_pstats;
    P.base = (double *)c->arena.base;
    P.chunk = chunk;
    P.op_cycles = nullptr;
    const char *prof_path = getenv("PMC_OP_PROFILE");  // diagnostic: per-operation cycle table of every launch
    if (prof_path && *prof_path) {
        CK(cudaMalloc((void **)&P.op_cycles, pg.ops.size() * 16));
        CK(cudaMemsetAsync(P.op_cycles, 0, pg.ops.size() * 16, c->stream));
    }
    EventPair ep;
    if (!c->ev_free.empty()) { ep = c->ev_free.back(); c->ev_free.pop_back(); }
    else { cudaEventCreate(&ep.a); cudaEventCreate(&ep.b); }
    cudaEventRecord(ep.a, c->stream);
    // CTA size: at most ~32 rows per row lane on the largest operand, enlarged while the tiles of the batch would not fill the
    // machine (nominal sizes; the kernel variants launched for them are listed below)
    int nt = 64;
    while (nt < 512 && max_rows / (nt / LPR) > 32) nt *= 2;
    while (nt < 512 && (long long)ntiles * nt * 2 <= (long long)c->num_sms * 1024) nt *= 2;
    // a batch that needs a second, mostly empty wave of CTAs runs better as one wave of smaller CTAs, as long as those
    // still fill at least half of the machine's threads (level 1 of the bench: 750 tiles, 20.5 -> 18.4 ms)
    while (c->single_wave && nt > 64 && (long long)ntiles * nt > (long long)c->num_sms * 1024 && (long long)ntiles * (nt / 2) <= (long long)c->num_sms * 1024 &&
           (long long)ntiles * (nt / 2) * 2 >= (long long)c->num_sms * 1024)
        nt /= 2;
    if (c->force_nt == 64 || c->force_nt == 128 || c->force_nt == 256 || c->force_nt == 512) nt = c->force_nt;
    // cluster size: split a tile over several CTAs while the batch has too few tiles to fill the machine and every
    // CTA keeps at least ~1000 rows of the largest operand
    int cs = 1;
    if (nt >= 256) {
        const long long slots = (long long)c->num_sms * (1024 / nt);
        while (cs < 8 && (long long)ntiles * cs * 2 <= slots && max_rows / (cs * 2) >= 1024) cs *= 2;
    }
    if (c->force_cs == 1 || ((c->force_cs == 2 || c->force_cs == 4 || c->force_cs == 8) && nt >= 256)) cs = c->force_cs;
    // grid groups: when even clusters of 8 would leave most of the machine idle (one or two tiles of a very large
    // level), a tile is split over G co-resident CTAs of a cooperative launch, every CTA keeping >= ~2000 rows
    int group = 0;
    {
        const long long slots = (long long)c->num_sms * 2;  // 512-thread CTAs, two per SM
        if (c->force_group > 1) group = c->force_group;
        else if (c->force_group == 0 && c->force_cs == 0 && nt == 512 && (long long)ntiles * 8 * 3 <= slots) {
            long long g = std::min<long long>(slots / ntiles, max_rows / 2048);
            if (g > 8) group = (int)g;
        }
        if (group > 1 && (long long)ntiles * group > slots) group = (int)(slots / ntiles);
        if (group <= 1) group = 0;
    }
    P.group = group;
    P.grp_bar = nullptr;
    P.grp_part = nullptr;
    if (group) {
        nt = 512;
        cs = 0;
        const size_t bar_bytes = (((size_t)ntiles * 2 * sizeof(unsigned int)) + 255) & ~(size_t)255;
        const size_t need = bar_bytes + (size_t)ntiles * group * TW * sizeof(double);
        if (need > c->grp_cap) {
            if (c->d_grp) { cudaStreamSynchronize(c->stream); cudaFree(c->d_grp); }
            c->d_grp = nullptr;
            c->grp_cap = 0;
            CK(cudaMalloc(&c->d_grp, need));
            c->grp_cap = need;
        }
        CK(cudaMemsetAsync(c->d_grp, 0, bar_bytes, c->stream));
        P.grp_bar = (unsigned int *)c->d_grp;
        P.grp_part = (double *)((char *)c->d_grp + bar_bytes);
        // operations that stay on the group's first CTA, and the barriers that can stay CTA-local
        std::vector<Op> &ops = pg.ops;
        auto solo_kind = [](int k) {
            return k == OP_SPMM || k == OP_CHEB_FIRST || k == OP_LINCOMB3 || k == OP_SOL_UPDATE || k == OP_SETUP_SPMM || k == OP_FILL ||
                   k == OP_COPY || k == OP_BROADCAST || k == OP_MAP_EXP || k == OP_CG_DIR;
        };
        auto scalar_kind = [](int k) {
            return k == OP_SC_INIT || k == OP_SC_ALPHA || k == OP_SC_BETA || k == OP_CHECK || k == OP_JUMP || k == OP_STORE_ITERS ||
                   k == OP_LIKELIHOOD || k == OP_CG_INIT || k == OP_CG_ALPHA || k == OP_CG_BETA || k == OP_CHB_INIT || k == OP_CHB_CHECK || k == OP_STORE_Q;
        };
        for (Op &o : ops) {
            o.flags &= ~(F_SOLO | F_LOCAL_SYNC);
            if (solo_kind(o.kind) && o.n <= c->solo_rows && !(o.flags & F_DOT)) o.flags |= F_SOLO;
            if (scalar_kind(o.kind)) o.flags |= F_LOCAL_SYNC;
        }
        for (size_t i = 0; i + 1 < ops.size(); ++i)
            if ((ops[i].flags & F_SOLO) && (ops[i + 1].flags & F_SOLO)) ops[i].flags |= F_LOCAL_SYNC;
    }
    memcpy(c->h_ops, pg.ops.data(), bytes);
    CK(cudaMemcpyAsync(c->d_ops, c->h_ops, bytes, cudaMemcpyHostToDevice, c->stream));
    cudaError_t le = cudaSuccess;
    // Kernel variants.  The nominal CTA sizes 512 / 256 / 128 / 64 of the selection above run as 448 / 224 / 128 / 64
    // threads with at most 896 (128 / 64: 768, measured 9 % faster on the coarsest level) threads resident per SM, which
    // leaves 72 (80) registers per thread: the interpreter with its
    // inlined sparse applies wants more than the 64 registers that 1024 resident threads would allow, and the better
    // schedule of the gather loops outweighs the lost eighth of the threads (level-0 batch of the bench: 54.2 -> 47.9 ms).
#ifndef PMC_NT_L
#define PMC_NT_L 448
#define PMC_NT_M 224
#define PMC_MINB_L 2
#define PMC_MINB_M 4
#define PMC_MINB_S 6
#define PMC_MINB_XS 12
#endif
#define PMC_LAUNCH(NT_, MINB_, CS_) le = launch_program<NT_, MINB_, CS_>(P, ntiles, c->stream)
    if (group) PMC_LAUNCH(PMC_NT_L, PMC_MINB_L, 0);
    else if (nt == 512) { if (cs == 8) PMC_LAUNCH(PMC_NT_L, PMC_MINB_L, 8); else if (cs == 4) PMC_LAUNCH(PMC_NT_L, PMC_MINB_L, 4); else if (cs == 2) PMC_LAUNCH(PMC_NT_L, PMC_MINB_L, 2); else PMC_LAUNCH(PMC_NT_L, PMC_MINB_L, 1); }
    else if (nt == 256) { if (cs == 8) PMC_LAUNCH(PMC_NT_M, PMC_MINB_M, 8); else if (cs == 4) PMC_LAUNCH(PMC_NT_M, PMC_MINB_M, 4); else if (cs == 2) PMC_LAUNCH(PMC_NT_M, PMC_MINB_M, 2); else PMC_LAUNCH(PMC_NT_M, PMC_MINB_M, 1); }
    else if (nt == 128) PMC_LAUNCH(128, PMC_MINB_S, 1);
    else PMC_LAUNCH(64, PMC_MINB_XS, 1);
#undef PMC_LAUNCH
    if (le != cudaSuccess && c->cuda_status == cudaSuccess) c->cuda_status = le;
    cudaEventRecord(ep.b, c->stream);
    if (P.op_cycles) {
        std::vector<unsigned long long> h(pg.ops.size() * 2);
        cudaStreamSynchronize(c->stream);
        cudaMemcpy(h.data(), P.op_cycles, h.size() * 8, cudaMemcpyDeviceToHost);
        cudaFree(P.op_cycles);
        if (FILE *f = fopen(prof_path, "a")) {
            fprintf(f, "# launch: %d tiles, %zu ops, nt %d cs %d\n# pc kind class n flags bytes_per_tile executions cycles_per_execution\n", ntiles, pg.ops.size(), nt, cs);
            for (size_t i = 0; i < pg.ops.size(); ++i)
                if (h[2 * i + 1])
                    fprintf(f, "%zu %d %d %d %d %.0f %llu %.0f\n", i, pg.ops[i].kind, pg.ops[i].kclass, pg.ops[i].n, pg.ops[i].flags,
                            pg.ops[i].bytes, h[2 * i + 1], (double)h[2 * i] / (double)h[2 * i + 1]);
            fclose(f);
        }
    }
    c->ev_pending.push_back(ep);
    c->kernel_launches++;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess && c->cuda_status == cudaSuccess) c->cuda_status = e;
    return PMC_OK;
}

static void resolve_events(Ctx *c)
{
    if (c->ev_pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto &ep : c->ev_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) c->kernel_ms += ms;
        c->ev_free.push_back(ep);
    }
    c->ev_pending.clear();
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

enum { CACHE_XI = 0, CACHE_FIELD = 1, CACHE_EMB = 2 };

// keep rows [s0, s0 + ns) of a result (n values per realisation, host layout, currently at `src` on the device)
static int cache_store(Ctx *c, int slot, int level, int nsamples, int s0, int ns, int n, const double *src)
{
    if (!c->cache_results) return PMC_OK;
    pmc_context_s::ResultCache &R = c->cache[slot];
    const size_t need = (size_t)nsamples * n;
    if (R.cap < need) {
        if (R.p) { cudaStreamSynchronize(c->stream); cudaFree(R.p); }
        R.p = nullptr; R.cap = 0; R.valid = false;
        CK(cudaMalloc((void **)&R.p, need * sizeof(double)));
        R.cap = need;
    }
    CK(cudaMemcpyAsync(R.p + (size_t)s0 * n, src, (size_t)ns * n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    R.level = level; R.ns = nsamples; R.n = n; R.valid = true;
    return PMC_OK;
}
static const double *cache_lookup(Ctx *c, int slot, int level, int nsamples, int n, const char *what)
{
    const pmc_context_s::ResultCache &R = c->cache[slot];
    if (!c->cache_results || !R.valid || R.level != level || R.ns != nsamples || R.n != n) {
        fail(c, PMC_ERR_STATE, "%s is NULL but the handle holds no matching device-resident result (option cache_results, "
                               "same level and number of realisations as the call that produced it)", what);
        return nullptr;
    }
    return R.p;
}

static dim3 grid1d(const Ctx *c, size_t n) { return dim3((unsigned)std::min<size_t>((n + 255) / 256, (size_t)c->num_sms * 16)); }

// host [ns][n] -> tile-major batched (via the staging buffer `stage`)
static int upload_rows(Ctx *c, const double *host, int ns, int n, double *stage, Off dst, Off chunk, int mode, double neg_g,
                       const double *w_sqrt, const int *rowmap = nullptr, const double *dev_src = nullptr)
{
    const int ntiles = (ns + TW - 1) / TW;
    double *d = (double *)c->arena.base + dst;
    if (dev_src) stage = const_cast<double *>(dev_src);   // already on the device in the host layout (result cache)
    else CK(cudaMemcpyAsync(stage, host, (size_t)ns * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const size_t total = (size_t)ntiles * n * TW;
    if (mode == 1) launch(c, PMC_K_MISC, (double)total * 16.0, k_to_tiles<1>, grid1d(c, total), dim3(256), n, chunk, ntiles, ns, (const double *)stage, d, neg_g, w_sqrt, rowmap);
    else launch(c, PMC_K_MISC, (double)total * 16.0, k_to_tiles<0>, grid1d(c, total), dim3(256), n, chunk, ntiles, ns, (const double *)stage, d, neg_g, w_sqrt, rowmap);
    return PMC_OK;
}

// tile-major batched view -> host [ns][n]
static int download_rows(Ctx *c, Off src_off, Off chunk, int ns, int n, double *stage, double *host, bool do_exp,
                         const int *rowmap = nullptr)
{
    const size_t total = (size_t)ns * n;
    const double *src = (const double *)c->arena.base + src_off;
    if (do_exp) launch(c, PMC_K_MISC, (double)total * 16.0, k_from_tiles<1>, grid1d(c, total), dim3(256), n, chunk, ns, src, stage, rowmap);
    else launch(c, PMC_K_MISC, (double)total * 16.0, k_from_tiles<0>, grid1d(c, total), dim3(256), n, chunk, ns, src, stage, rowmap);
    CK(cudaMemcpyAsync(host, stage, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
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
    c->num_sms = prop.multiProcessorCount;
    c->store = std::make_shared<DevStore>();
    c->store->device = device;
    c->cfg_sampler.max_vlevels = -1;  // single-level Schur smoother when alpha*W dominates (short correlation length)
    c->cfg_darcy.omega = 2.5;         // aggregation-type coarse spaces under-correct the pressure Laplacian
    c->cfg_darcy.mass_degree = 1;     // plain Jacobi on the RT mass block: fewest bytes per unit of convergence
    c->cfg_sampler.mass_degree = 1;   // (measured on the bench hierarchy, tools/tune_prec.py)
    c->s.resize(nlevels);
    c->d.resize(nlevels);
    memset(&c->stats, 0, sizeof c->stats);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&c->d_pstats, sizeof(ProgStats)) != cudaSuccess ||
        cudaMalloc((void **)&c->d_tab, sizeof(RngTables)) != cudaSuccess) {
        delete c;
        return fail(nullptr, PMC_ERR_CUDA, "pmc_create: CUDA resource allocation failed");
    }
    cudaMemset(c->d_pstats, 0, sizeof(ProgStats));
    *out = c;
    return PMC_OK;
}

void pmc_destroy(pmc_handle c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->nccl_comm) pmc_comm_destroy(c);
    if (c->d_comm_buf) cudaFree(c->d_comm_buf);
    for (auto &R : c->cache) if (R.p) cudaFree(R.p);
    c->store.reset();   // frees the operators with the last handle that shares them
    if (c->arena.base) cudaFree(c->arena.base);
    cudaFree(c->d_pstats);
    cudaFree(c->d_tab);
    if (c->d_ops) cudaFree(c->d_ops);
    if (c->d_grp) cudaFree(c->d_grp);
    if (c->h_ops) cudaFreeHost(c->h_ops);
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
    CK(cudaSetDevice(c->device));
    resolve_events(c);
    cudaStreamSynchronize(c->stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return PMC_OK;
}

int pmc_synchronize(pmc_handle c)
{
    if (!c) return PMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
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
        if (coarse_degree > 0) { g->coarse_degree = coarse_degree; g->coarse_user = true; }
        if (coarse_ratio > 1.0) { g->coarse_ratio = coarse_ratio; g->coarse_user = true; }
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
        else if (k == "schur_degree_coarse" && value >= -1) g->schur_degree_coarse = (int)value;
        else if (k == "schur_ratio" && value > 1) g->schur_ratio = value;
        else if (k == "coarse_degree" && value >= 1) { g->coarse_degree = (int)value; g->coarse_user = true; }
        else if (k == "coarse_ratio" && value > 1) { g->coarse_ratio = value; g->coarse_user = true; }
        else if (k == "omega" && value > 0) { g->omega = value; g->omega_user = true; }
        else if (k == "mass_scale" && value > 0) g->mass_scale = value;
        else if (k == "cheb_lo_scale" && value > 0) g->cheb_lo_scale = value;
        else if (k == "cheb_hi_scale" && value > 0) g->cheb_hi_scale = value;
        else if (k == "max_vlevels") g->max_vlevels = (int)value;
        else if (k == "method") g->method = (int)value;
        else if (k == "amg") g->amg = (int)value;
        else if (k == "amg_smooth" && value >= 0) g->amg_smooth = value;
        else if (k == "p_smooth" && value >= 0) g->p_smooth = value;
        else if (k == "amg_passes" && value >= 1 && value <= 6) g->amg_passes = (int)value;
        else return fail(c, PMC_ERR_ARG, "pmc_set_option: bad key or value '%s' = %g", key, value);
        return PMC_OK;
    }
    if (k == "cache_results") { c->cache_results = value != 0; return PMC_OK; }
    if (k == "split_apply") { c->split_apply = value != 0; return PMC_OK; }
    if (k == "stage_wide") { c->stage_wide = value != 0; return PMC_OK; }
    if (k == "max_batch" && value >= 0) c->max_batch = (int)value;
    else if (k == "cta_threads") c->force_nt = (int)value;
    else if (k == "cluster_size") c->force_cs = (int)value;
    else if (k == "stage_operators") c->staging = value != 0;
    else if (k == "defer_x") c->defer_x = value != 0;
    else if (k == "cheb_three_term") c->cheb3 = value != 0;
    else if (k == "qoi_only") c->qoi_only = value != 0;
    else if (k == "fuse_coarse") c->fuse_coarse = value != 0;
    else if (k == "single_wave") c->single_wave = value != 0;
    else if (k == "renumber") {
        for (int l = 0; l < c->nlevels; ++l)
            if (c->s[l].set || c->d[l].set) return fail(c, PMC_ERR_STATE, "pmc_set_option(renumber): set it before the uploads");
        c->renumber = value != 0;
    }
    else if (k == "group_size") c->force_group = (int)value;
    else if (k == "solo_rows" && value >= 0) c->solo_rows = (int)value;
    else return fail(c, PMC_ERR_ARG, "pmc_set_option: unknown key '%s'", key);
    return PMC_OK;
}

int pmc_set_batch(pmc_handle c, int max_batch, int check_every)
{
    if (!c) return PMC_ERR_ARG;
    if (max_batch >= 0) c->max_batch = max_batch;
    (void)check_every;  // convergence is checked on the device after every iteration; kept for ABI stability
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
    if (c->renumber) {  // the u block never crosses the ABI for the sampler: renumber and forget
        const std::vector<int> perm = first_touch_order(Nf, Ne, L.B.rowptr.data(), L.B.col.data());
        L.M = csr_permuted(L.M, perm.data(), perm.data());
        L.B = csr_permuted(L.B, nullptr, perm.data());
    }
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
    if (c->renumber) {
        const std::vector<int> perm = first_touch_order(Nf, Ne, L.elem_ptr.data(), L.elem_dofs.data());
        for (int &d : L.elem_dofs) d = perm[d];
        L.B = csr_permuted(L.B, nullptr, perm.data());
        permute_head(L.ess_u, perm);
        permute_head(L.ess_data, perm);
        permute_head(L.rhs, perm);
        permute_head(L.obs, perm);
        std::vector<int> rowmap(Nf + Ne);
        for (int f = 0; f < Nf; ++f) rowmap[f] = perm[f];
        for (int e = 0; e < Ne; ++e) rowmap[Nf + e] = Nf + e;
        int rc = to_device(c, rowmap, &L.d_rowmap);
        if (rc) return rc;
        L.h_rowmap = std::move(rowmap);
    }
    L.set = true;
    return PMC_OK;
}

int pmc_upload_field_transfer(pmc_handle c, int level, int n_out, const int *T_rowptr, const int *T_col, const double *T_val,
                              const double *row_scale)
{
    if (!c) return PMC_ERR_ARG;
    if (level < 0 || level >= c->nlevels || n_out < 1 || !T_rowptr || !T_col || !T_val)
        return fail(c, PMC_ERR_ARG, "pmc_upload_field_transfer: bad arguments");
    SamplerLevel &L = c->s[level];
    if (!L.set) return fail(c, PMC_ERR_STATE, "pmc_upload_field_transfer: sampler level %d not uploaded", level);
    if (L.hasT) return fail(c, PMC_ERR_STATE, "field transfer of level %d uploaded twice", level);
    CK(cudaSetDevice(c->device));
    L.T = csr_copy(n_out, L.Ne, T_rowptr, T_col, T_val);
    for (int i = 0; i < n_out; ++i)
        for (int p = L.T.rowptr[i]; p < L.T.rowptr[i + 1]; ++p) {
            if (L.T.col[p] < 0 || L.T.col[p] >= L.Ne) return fail(c, PMC_ERR_ARG, "field transfer: column out of range");
            if (row_scale) L.T.val[p] *= row_scale[i];
        }
    int rc = upload_csr(c, L.T, L.dT);
    if (rc) return rc;
    L.n_out = n_out;
    L.hasT = true;
    return PMC_OK;
}

int pmc_host_alloc(size_t bytes, void **out)
{
    if (!out) return PMC_ERR_ARG;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, PMC_ERR_NOMEM, "pmc_host_alloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return PMC_OK;
}

void pmc_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int pmc_clone(pmc_handle src, pmc_handle *out)
{
    if (!src || !out) return PMC_ERR_ARG;
    *out = nullptr;
    // The clone shares the source's device operators (DevStore): build every derived structure of the source first, so
    // that the copied level descriptors are complete and the clone never has to prepare anything itself.
    int rc = pmc_prepare(src);
    if (rc) return rc;
    pmc_handle c = nullptr;
    rc = pmc_create(src->device, src->nlevels, &c);
    if (rc) return fail(src, rc, "pmc_clone: %s", pmc_last_error(nullptr));
    c->rel = src->rel; c->abs_ = src->abs_; c->maxit = src->maxit;
    c->cfg_sampler = src->cfg_sampler; c->cfg_darcy = src->cfg_darcy;
    c->max_batch = src->max_batch; c->force_nt = src->force_nt; c->force_cs = src->force_cs; c->staging = src->staging;
    c->defer_x = src->defer_x; c->cheb3 = src->cheb3; c->qoi_only = src->qoi_only; c->fuse_coarse = src->fuse_coarse; c->single_wave = src->single_wave; c->renumber = src->renumber;
    c->force_group = src->force_group; c->solo_rows = src->solo_rows; c->cache_results = src->cache_results; c->split_apply = src->split_apply; c->stage_wide = src->stage_wide;
    c->store = src->store;   // one copy of the operators per device, freed with the last handle that uses them
    c->s = src->s;           // level descriptors: host arrays by value, device pointers into the shared store
    c->d = src->d;
    if (src->rng_ready) rc = pmc_rng_init(c, src->mu, src->sigma, src->rng_nparts, src->rng_mypart);
    if (rc) {
        fail(src, rc, "pmc_clone: %s", pmc_last_error(c));
        pmc_destroy(c);
        return rc;
    }
    *out = c;
    return PMC_OK;
}

// ---- ranks that share a sample budget: NCCL all-reduce of the per-level sums ------------------------------------------
// (/root/reference/src/MLMC_Manager.cpp computes every rank's samples redundantly and needs no reduction; here the
// ranks own disjoint slices of every level's budget, SURVEY section 8e, so InitRun ends with one all-reduce.)
// libnccl is loaded at the first use (dlopen), so the library itself links only the CUDA runtime and loads on hosts
// without NCCL; a process that already holds libnccl.so.2 (torch) shares that copy.
namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char *names[] = {getenv("PMC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.lib) break;
            api.err = dlerror();
        }
        if (!api.lib) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy || !api.GetErrorString) {
            api.err = "libnccl lacks a required symbol";
            api.lib = nullptr;
        }
    });
    return api.lib ? &api : nullptr;
}
const char *nccl_load_error()
{
    static NcclApi *probe = nccl_api();
    (void)probe;
    return "libnccl.so.2 could not be loaded (set PMC_NCCL_LIB)";
}
}  // namespace

static_assert(PMC_COMM_ID_BYTES == sizeof(ncclUniqueId), "PMC_COMM_ID_BYTES must be the size of ncclUniqueId");

int pmc_comm_unique_id(void *id_out)
{
    if (!id_out) return fail(nullptr, PMC_ERR_ARG, "pmc_comm_unique_id: null argument");
    NcclApi *n = nccl_api();
    if (!n) return fail(nullptr, PMC_ERR_STATE, "pmc_comm_unique_id: %s", nccl_load_error());
    ncclUniqueId id;
    ncclResult_t r = n->GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, PMC_ERR_CUDA, "ncclGetUniqueId: %s", n->GetErrorString(r));
    memcpy(id_out, &id, sizeof id);
    return PMC_OK;
}

int pmc_comm_init(pmc_handle c, int nranks, int rank, const void *id)
{
    if (!c) return PMC_ERR_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, PMC_ERR_ARG, "pmc_comm_init: rank %d of %d", rank, nranks);
    if (c->nccl_comm) pmc_comm_destroy(c);
    c->comm_ranks = nranks;
    c->comm_rank = rank;
    if (nranks == 1) return PMC_OK;
    if (!id) return fail(c, PMC_ERR_ARG, "pmc_comm_init: the unique id is required for more than one rank");
    NcclApi *n = nccl_api();
    if (!n) return fail(c, PMC_ERR_STATE, "pmc_comm_init: %s", nccl_load_error());
    CK(cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t comm = nullptr;
    ncclResult_t r = n->CommInitRank(&comm, nranks, uid, rank);
    if (r != ncclSuccess) return fail(c, PMC_ERR_CUDA, "ncclCommInitRank: %s", n->GetErrorString(r));
    c->nccl_comm = comm;
    return PMC_OK;
}

int pmc_comm_destroy(pmc_handle c)
{
    if (!c) return PMC_ERR_ARG;
    if (c->nccl_comm) {
        NcclApi *n = nccl_api();
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (n) n->CommDestroy((ncclComm_t)c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    c->comm_ranks = 1;
    c->comm_rank = 0;
    return PMC_OK;
}

int pmc_allreduce_sums(pmc_handle c, double *sums, int count)
{
    if (!c || count < 0 || (count > 0 && !sums)) return PMC_ERR_ARG;
    if (c->comm_ranks == 1 || count == 0) return PMC_OK;
    if (!c->nccl_comm) return fail(c, PMC_ERR_STATE, "pmc_allreduce_sums: pmc_comm_init has not been called");
    NcclApi *n = nccl_api();
    CK(cudaSetDevice(c->device));
    if (c->comm_buf_count < (size_t)count) {
        if (c->d_comm_buf) cudaFree(c->d_comm_buf);
        c->d_comm_buf = nullptr;
        c->comm_buf_count = 0;
        CK(cudaMalloc((void **)&c->d_comm_buf, (size_t)count * sizeof(double)));
        c->comm_buf_count = (size_t)count;
    }
    int rc = ensure_pinned(c, (size_t)count);
    if (rc) return rc;
    memcpy(c->h_pinned, sums, (size_t)count * sizeof(double));
    CK(cudaMemcpyAsync(c->d_comm_buf, c->h_pinned, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ncclResult_t r = n->AllReduce(c->d_comm_buf, c->d_comm_buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)c->nccl_comm, c->stream);
    if (r != ncclSuccess) return fail(c, PMC_ERR_CUDA, "ncclAllReduce: %s", n->GetErrorString(r));
    CK(cudaMemcpyAsync(c->h_pinned, c->d_comm_buf, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    memcpy(sums, c->h_pinned, (size_t)count * sizeof(double));
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
    CK(cudaStreamSynchronize(c->stream));
    RngTables *t = new RngTables();
    h_build_rng_tables(*t, nparts, mypart);
    cudaError_t e = cudaMemcpy(c->d_tab, t, sizeof(RngTables), cudaMemcpyHostToDevice);
    delete t;
    if (e != cudaSuccess) return fail(c, PMC_ERR_CUDA, "rng table upload failed: %s", cudaGetErrorString(e));
    c->mu = mu;
    c->sigma = sigma;
    c->rng_nparts = nparts;
    c->rng_mypart = mypart;
    c->rng_ready = true;
    return PMC_OK;
}

static int rng_fill_common(Ctx *c, int mode, uint64_t pos, int64_t n, void *out, int cache_level = -1, int cache_ns = 0)
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
    if (cache_level >= 0 && cache_ns > 0 &&
        (rc = cache_store(c, CACHE_XI, cache_level, cache_ns, 0, cache_ns, (int)(n / cache_ns), (const double *)c->arena.base)))
        return rc;
    CK(cudaMemcpyAsync(out, c->arena.base, (size_t)n * esz, cudaMemcpyDeviceToHost, c->stream));
    return finish(c);
}

int pmc_rng_fill_int(pmc_handle c, uint64_t pos, int64_t n, int32_t *out) { return rng_fill_common(c, 0, pos, n, out); }
int pmc_rng_fill(pmc_handle c, uint64_t pos, int64_t n, double *out) { return rng_fill_common(c, 1, pos, n, out); }

int pmc_rng_map(pmc_handle c, int64_t n, const int32_t *engine, double *out)
{
    if (!c || n < 0 || (n > 0 && (!out || !engine))) return PMC_ERR_ARG;
    if (!c->rng_ready) return fail(c, PMC_ERR_STATE, "pmc_rng_init has not been called");
    if (n == 0) return PMC_OK;
    CK(cudaSetDevice(c->device));
    const size_t off = ((size_t)n * 8 + 255) & ~(size_t)255;
    int rc = ensure_arena(c, off + (size_t)n * 4 + 4096);
    if (rc) return rc;
    double *d_out = (double *)c->arena.base;
    int32_t *d_in = (int32_t *)((char *)c->arena.base + off);
    CK(cudaMemcpyAsync(d_in, engine, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    launch(c, PMC_K_RNG, (double)n * 12.0, k_rng_map, grid1d(c, (size_t)n), dim3(256), n, (const int32_t *)d_in, d_out, c->mu, c->sigma);
    CK(cudaMemcpyAsync(out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    return finish(c);
}

int pmc_sampler_sample_batch(pmc_handle c, int level, int nsamples, uint64_t pos0, double *xi_out)
{
    if (!c) return PMC_ERR_ARG;
    if (level < 0 || level >= c->nlevels || !c->s[level].set) return fail(c, PMC_ERR_STATE, "sampler level %d not uploaded", level);
    // realisation j occupies positions pos0 + j*Ne ... : a flat fill of nsamples*Ne values
    return rng_fill_common(c, 1, pos0, (int64_t)nsamples * c->s[level].Ne, xi_out, level, nsamples);
}

// ---- sampler Eval ---------------------------------------------------------------------------------
// Restrict a batched right-hand side from xi_level to level (Ps^T chain, /root/reference/src/PDESampler.cpp:361-368);
// returns the buffer holding the result.
static Off emit_restrict(Program &pg, Ctx *c, int xi_level, int level, Off cur, Off other)
{
    for (int l = xi_level; l < level; ++l) {
        SamplerLevel &L = c->s[l];
        const int nc = c->s[l + 1].Ne;
        emit_spmm(pg, KC_TRANSFER, EP_AX, L.dPt, VNULL, vr(cur, L.Ne), vr(other, nc), VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false,
                  false, (double)L.Ne + nc);
        std::swap(cur, other);
    }
    return cur;
}

// Prolongate a batched Gaussian field from init_level to the finer `level` (:496-508).
static Off emit_prolong(Program &pg, Ctx *c, int init_level, int level, Off cur, Off other)
{
    for (int l = init_level; l > level; --l) {
        SamplerLevel &L = c->s[l - 1];
        const int nc = c->s[l].Ne;
        emit_spmm(pg, KC_TRANSFER, EP_AX, L.dP, VNULL, vr(cur, nc), vr(other, L.Ne), VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false,
                  false, (double)L.Ne + nc);
        std::swap(cur, other);
    }
    return cur;
}

int pmc_sampler_eval_batch(pmc_handle c, int level, int xi_level, int nsamples, const double *xi, const double *init_s,
                           int init_level, int use_init, double *s_out, double *embed_s_out, int *iters_out)
{
    int rc = check_level(c, level, true, false);
    if (rc) return rc;
    if (nsamples < 0 || !s_out) return fail(c, PMC_ERR_ARG, "pmc_sampler_eval_batch: bad arguments");
    if (xi_level < 0 || xi_level > level || !c->s[xi_level].set)
        return fail(c, PMC_ERR_ARG, "xi_level %d must be an uploaded level <= level %d", xi_level, level);
    const bool warm = use_init > 0 && (init_s != nullptr || (c->cache_results && c->cache[CACHE_EMB].valid));
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
    const double *xi_dev = nullptr, *init_dev = nullptr;
    if (!xi && !(xi_dev = cache_lookup(c, CACHE_XI, xi_level, nsamples, Nex, "xi"))) return PMC_ERR_STATE;
    if (warm && !init_s && !(init_dev = cache_lookup(c, CACHE_EMB, init_level, nsamples, Nei, "init_s"))) return PMC_ERR_STATE;
    // the program (identical for every batch): restrict, optional prolongated initial guess, solve
    Rows ar;
    SolveWs ws;
    carve_solve(ar, sys, ws);
    const Off bufA = ar.alloc(nmax), bufB = ar.alloc(nmax), bufC = ar.alloc(nmax);
    Program pg;
    pg.staging = c->staging;
    pg.defer_x = c->defer_x;
    pg.cheb3 = c->cheb3;
    pg.qoi_only = c->qoi_only;
    pg.fuse_coarse = c->fuse_coarse;
    pg.split_apply = c->split_apply;
    pg.chunked = c->stage_wide;
    const Off rhs = emit_restrict(pg, c, xi_level, level, bufA, bufB);
    const Off t1 = (rhs == bufA) ? bufB : bufA;
    Off x0 = -1;
    if (warm) x0 = emit_prolong(pg, c, init_level, level, t1, bufC);
    emit_sampler_solve(pg, c, level, rhs, x0, ws, true);
    const int Nout = L.out_size();
    Off fout = ws.x + (Off)sys.Nf * TW;  // rows [Nf, N) of the solution: the Gaussian field
    if (L.hasT) {
        // project to the forward problem's mesh before exp (/root/reference/src/EmbeddedPDESampler.cpp:426-435,
        // src/L2ProjectionPDESampler.cpp:595-611)
        const Off buf = ar.alloc(Nout);
        emit_spmm(pg, KC_TRANSFER, EP_AX, L.dT, VNULL, vr(fout), vr(buf), VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false,
                  (double)Ne + Nout);
        fout = buf;
    }
    const Off chunk = ar.peak;
    const size_t per_sample = (size_t)chunk * 8 / TW + (size_t)std::max(nmax, Nout) * 8 + 64;
    int B = 0;
    if ((rc = size_batch(c, per_sample, 0, nsamples, &B))) return rc;
    std::vector<double> itbuf;
    for (int s0 = 0; s0 < nsamples; s0 += B) {
        const int ns = std::min(B, nsamples - s0);
        const int ld = pad_ld(ns);
        double *stage = (double *)c->arena.base + (size_t)(ld / TW) * (size_t)chunk;
        // rhs_s = -g * xi * w_sqrt at xi_level (:352-358 / :423-428)
        if ((rc = upload_rows(c, xi ? xi + (size_t)s0 * Nex : nullptr, ns, Nex, stage, bufA, chunk, 1, -c->s[xi_level].g,
                              c->s[xi_level].w_sqrt, nullptr, xi_dev ? xi_dev + (size_t)s0 * Nex : nullptr))) return rc;
        if (warm && (rc = upload_rows(c, init_s ? init_s + (size_t)s0 * Nei : nullptr, ns, Nei, stage, t1, chunk, 0, 0.0, nullptr,
                                      nullptr, init_dev ? init_dev + (size_t)s0 * Nei : nullptr))) return rc;
        if ((rc = run_program(c, pg, ns, chunk, sys.N))) return rc;
        const Off field = ws.x + (Off)sys.Nf * TW;  // rows [Nf, N) of the solution
        if ((rc = download_rows(c, fout, chunk, ns, Nout, stage, s_out + (size_t)s0 * Nout, L.lognormal != 0))) return rc;
        if ((rc = cache_store(c, CACHE_FIELD, level, nsamples, s0, ns, Nout, stage))) return rc;
        if (embed_s_out) {
            if ((rc = download_rows(c, field, chunk, ns, Ne, stage, embed_s_out + (size_t)s0 * Ne, false))) return rc;
            if ((rc = cache_store(c, CACHE_EMB, level, nsamples, s0, ns, Ne, stage))) return rc;
        }
        if (iters_out) {
            itbuf.resize(ns);
            if ((rc = download_rows(c, ws.iters, chunk, ns, 1, stage, itbuf.data(), false))) return rc;
        }
        if ((rc = finish(c))) return rc;
        if (iters_out)
            for (int j = 0; j < ns; ++j) iters_out[s0 + j] = (int)itbuf[j];
    }
    return PMC_OK;
}

// ---- Darcy ----------------------------------------------------------------------------------------
static int darcy_host_batch(Ctx *c, int level, int nsamples, const double *k, const double *xin, double *Q_out,
                            double *C_out, double *sol_out, int *iters_out, bool apply_only)
{
    int rc = check_level(c, level, false, true);
    if (rc) return rc;
    if (nsamples < 0) return fail(c, PMC_ERR_ARG, "Darcy batch: bad arguments");
    if (nsamples == 0) return PMC_OK;
    CK(cudaSetDevice(c->device));
    DarcyLevel &L = c->d[level];
    SaddleSys &sys = L.sys;
    const int Ne = L.Ne, N = sys.N;
    Rows ar;
    SolveWs ws;
    carve_solve(ar, sys, ws);
    const Off k_ext = ar.alloc(Ne + 1), Qrow = ar.alloc(1);
    const double *k_dev = nullptr;
    if (!k && !(k_dev = cache_lookup(c, CACHE_FIELD, level, nsamples, Ne, "k"))) return PMC_ERR_STATE;
    Program pg;
    pg.staging = c->staging;
    pg.defer_x = c->defer_x;
    pg.cheb3 = c->cheb3;
    pg.qoi_only = c->qoi_only;
    pg.fuse_coarse = c->fuse_coarse;
    pg.split_apply = c->split_apply;
    pg.chunked = c->stage_wide;
    if (apply_only) {
        Solver sv{&sys, &ws, vr(k_ext, Ne + 1)};
        emit_fill(pg, vr(k_ext, Ne + 1, Ne), 1, 1.0);  // weight of the fixed entries
        emit_saddle(pg, sv, EP_AX, ws.x, ws.q, -1, -1);
    } else {
        emit_fill(pg, vr(k_ext, Ne + 1, Ne), 1, 1.0);
        emit_darcy_solve(pg, c, level, k_ext, ws, Qrow, true, sol_out != nullptr);
    }
    const Off chunk = ar.peak;
    const size_t per_sample = (size_t)chunk * 8 / TW + (size_t)N * 8 + 64;
    int B = 0;
    if ((rc = size_batch(c, per_sample, 0, nsamples, &B))) return rc;
    std::vector<double> itbuf;
    for (int s0 = 0; s0 < nsamples; s0 += B) {
        const int ns = std::min(B, nsamples - s0);
        const int ld = pad_ld(ns);
        double *stage = (double *)c->arena.base + (size_t)(ld / TW) * (size_t)chunk;
        if ((rc = upload_rows(c, k ? k + (size_t)s0 * Ne : nullptr, ns, Ne, stage, k_ext, chunk, 0, 0.0, nullptr, nullptr,
                              k_dev ? k_dev + (size_t)s0 * Ne : nullptr))) return rc;
        if (apply_only) {
            if ((rc = upload_rows(c, xin + (size_t)s0 * N, ns, N, stage, ws.x, chunk, 0, 0.0, nullptr, L.d_rowmap))) return rc;
            if ((rc = run_program(c, pg, ns, chunk, N))) return rc;
            if ((rc = download_rows(c, ws.q, chunk, ns, N, stage, sol_out + (size_t)s0 * N, false, L.d_rowmap))) return rc;
            if ((rc = finish(c))) return rc;
            continue;
        }
        if ((rc = run_program(c, pg, ns, chunk, N))) return rc;
        if (Q_out && (rc = download_rows(c, Qrow, chunk, ns, 1, stage, Q_out + s0, false))) return rc;
        if (C_out)
            for (int j = 0; j < ns; ++j) C_out[s0 + j] = (double)N;  // /root/reference/src/DarcySolver.cpp:429
        if (sol_out && (rc = download_rows(c, ws.x, chunk, ns, N, stage, sol_out + (size_t)s0 * N, false, L.d_rowmap))) return rc;
        if (iters_out) {
            itbuf.resize(ns);
            if ((rc = download_rows(c, ws.iters, chunk, ns, 1, stage, itbuf.data(), false))) return rc;
        }
        if ((rc = finish(c))) return rc;
        if (iters_out)
            for (int j = 0; j < ns; ++j) iters_out[s0 + j] = (int)itbuf[j];
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
// One level of the manager loop as ONE program / one kernel launch per batch.
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
    const int Nk = SF.out_size(), Nkc = coarsest ? 0 : c->s[level + 1].out_size();  // size of the coefficient vectors
    if (Nk != DF.Ne || (!coarsest && Nkc != c->d[level + 1].Ne))
        return fail(c, PMC_ERR_STATE, "sampler output size and Darcy coefficient size differ on level %d (missing field transfer?)", level);
    const double cost = (double)DF.sys.N + (coarsest ? 0.0 : (double)c->d[level + 1].sys.N);
    // ---- record the program of the level (the same for every batch except the stream position) ----
    Rows ar;
    const Off rhs_f = ar.alloc(Ne);
    const Off rhs_c = coarsest ? -1 : ar.alloc(Nec), s_c = coarsest ? -1 : ar.alloc(Nec), x0_f = coarsest ? -1 : ar.alloc(Ne);
    const Off k_ext = ar.alloc(std::max(Nk, Nkc) + 1), Qf = ar.alloc(1), Qc = ar.alloc(1);
    const Off tbuf = (SF.hasT || (!coarsest && c->s[level + 1].hasT)) ? ar.alloc(std::max(Nk, Nkc)) : -1;
    const Off mark = ar.top;
    Program pg;
    pg.staging = c->staging;
    pg.defer_x = c->defer_x;
    pg.cheb3 = c->cheb3;
    pg.qoi_only = c->qoi_only;
    pg.fuse_coarse = c->fuse_coarse;
    pg.split_apply = c->split_apply;
    pg.chunked = c->stage_wide;
    // Sample(level, xi) fused with rhs_s = -g W^{1/2} xi  (/root/reference/src/PDESampler.cpp:336-340, :352-358)
    {
        Op &o = pg.add(OP_RNG, KC_RNG, Ne, Ne);
        o.y = vr(rhs_f, Ne);
        o.fixed = SF.w_sqrt;
        o.ca = -SF.g;
        o.a0 = 1;
    }
    const int rng_op = 0;
    if (!coarsest) {
        SamplerLevel &SC = c->s[level + 1];
        DarcyLevel &DC = c->d[level + 1];
        // Eval(level+1, xi, ., init_s, false): restrict the right-hand side, solve from zero (:431-438)
        emit_spmm(pg, KC_TRANSFER, EP_AX, SF.dPt, VNULL, vr(rhs_f, Ne), vr(rhs_c, Nec), VNULL, VNULL, nullptr, VNULL, 0, 0, -1,
                  false, false, (double)Ne + Nec);
        {
            ar.top = mark;
            SolveWs ws;
            carve_solve(ar, SC.sys, ws);
            emit_sampler_solve(pg, c, level + 1, rhs_c, -1, ws, false);
            emit_copy(pg, vr(ws.x, SC.sys.N, SC.sys.Nf), vr(s_c, Nec), Nec);
            Off fsrc = s_c;
            if (SC.hasT) {  // to the forward problem's mesh, before exp
                emit_spmm(pg, KC_TRANSFER, EP_AX, SC.dT, VNULL, vr(s_c), vr(tbuf), VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false,
                          false, (double)Nec + Nkc);
                fsrc = tbuf;
            }
            if (SC.lognormal) { Op &o = pg.add(OP_MAP_EXP, KC_MISC, Nkc, 2.0 * Nkc); o.x = vr(fsrc); o.y = vr(k_ext); }
            else emit_copy(pg, vr(fsrc), vr(k_ext), Nkc);
            emit_fill(pg, vr(k_ext, 0, Nkc), 1, 1.0);
        }
        {
            ar.top = mark;
            SolveWs ws;
            carve_solve(ar, DC.sys, ws);
            emit_darcy_solve(pg, c, level + 1, k_ext, ws, Qc, false, false);
        }
        // initial guess for the fine solve: prolongated coarse Gaussian field (:496-511)
        emit_spmm(pg, KC_TRANSFER, EP_AX, SF.dP, VNULL, vr(s_c, Nec), vr(x0_f, Ne), VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false,
                  false, (double)Ne + Nec);
    }
    {
        ar.top = mark;
        SolveWs ws;
        carve_solve(ar, SF.sys, ws);
        emit_sampler_solve(pg, c, level, rhs_f, x0_f, ws, false);
        VecRef fsrc = vr(ws.x, SF.sys.N, SF.sys.Nf);
        if (SF.hasT) {
            emit_spmm(pg, KC_TRANSFER, EP_AX, SF.dT, VNULL, fsrc, vr(tbuf), VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false,
                      (double)Ne + Nk);
            fsrc = vr(tbuf);
        }
        if (SF.lognormal) { Op &o = pg.add(OP_MAP_EXP, KC_MISC, Nk, 2.0 * Nk); o.x = fsrc; o.y = vr(k_ext); }
        else emit_copy(pg, fsrc, vr(k_ext), Nk);
        emit_fill(pg, vr(k_ext, 0, Nk), 1, 1.0);
    }
    {
        ar.top = mark;
        SolveWs ws;
        carve_solve(ar, DF.sys, ws);
        emit_darcy_solve(pg, c, level, k_ext, ws, Qf, false, false);
    }
    const Off chunk = ar.peak;
    const size_t per_sample = (size_t)chunk * 8 / TW + 64 + 40;
    int B = 0;
    if ((rc = size_batch(c, per_sample, 32, nsamples, &B))) return rc;
    if ((rc = ensure_pinned(c, 16 + (rows ? (size_t)B * 4 : 0)))) return rc;
    // iterations of THIS call: the device counter is read (stream-ordered) before the first launch, because host-buffer
    // calls made on the handle in between advance it without updating iters_seen
    unsigned long long it0 = c->iters_seen;
    for (int s0 = 0; s0 < nsamples; s0 += B) {
        const int ns = std::min(B, nsamples - s0);
        const int ld = pad_ld(ns);
        double *out9 = (double *)c->arena.base + (size_t)(ld / TW) * (size_t)chunk;
        double *rows_d = rows ? out9 + 16 : nullptr;
        pg.ops[rng_op].u0 = pos0 + (uint64_t)s0 * (uint64_t)Ne;
        if (s0 == 0)
            CK(cudaMemcpyAsync(c->h_pinned + 11, &c->d_pstats->iters_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        if ((rc = run_program(c, pg, ns, chunk, DF.sys.N))) return rc;
        launch(c, PMC_K_MISC, (double)ns * 16.0, k_mlmc_accumulate, dim3(1), dim3(256), ns, (const double *)c->arena.base, chunk,
               Qf, coarsest ? (Off)-1 : Qc, cost, out9, rows_d);
        CK(cudaMemcpyAsync(c->h_pinned, out9, 9 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(c->h_pinned + 10, &c->d_pstats->iters_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        if (rows) CK(cudaMemcpyAsync(c->h_pinned + 16, rows_d, (size_t)ns * 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if ((rc = finish(c))) return rc;
        const double *o = c->h_pinned;
        if (s0 == 0) memcpy(&it0, c->h_pinned + 11, sizeof(unsigned long long));
        memcpy(&c->iters_seen, c->h_pinned + 10, sizeof(unsigned long long));
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
    if (total_iters) *total_iters = (int64_t)(c->iters_seen - it0);
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

// ---- Bayesian inverse problem ----------------------------------------------------------------------
int pmc_upload_observations(pmc_handle c, int level, int m, const double *g, const double *G_obs, double noise)
{
    if (!c) return PMC_ERR_ARG;
    if (level < 0 || level >= c->nlevels || m < 1 || !g || !G_obs || !(noise > 0.0))
        return fail(c, PMC_ERR_ARG, "pmc_upload_observations: bad arguments");
    DarcyLevel &L = c->d[level];
    if (!L.set) return fail(c, PMC_ERR_STATE, "pmc_upload_observations: Darcy level %d not uploaded", level);
    CK(cudaSetDevice(c->device));
    L.h_gobs_func.assign((size_t)m * L.Ne, 0.0);
    for (int i = 0; i < m; ++i) {
        double sum = 0.0;  // ComputeG divides by sum(g_obs_func[i]) (/root/reference/src/BayesianInverseProblem.cpp:186-187)
        for (int e = 0; e < L.Ne; ++e) sum += g[(size_t)i * L.Ne + e];
        if (sum == 0.0) return fail(c, PMC_ERR_ARG, "observation functional %d sums to zero on level %d", i, level);
        for (int e = 0; e < L.Ne; ++e) L.h_gobs_func[(size_t)i * L.Ne + e] = g[(size_t)i * L.Ne + e] / sum;
    }
    L.h_Gobs.assign(G_obs, G_obs + m);
    L.n_obs = m;
    L.noise = noise;
    int rc;
    if ((rc = to_device(c, L.h_gobs_func, &L.d_gobs_func))) return rc;
    if ((rc = to_device(c, L.h_Gobs, &L.d_Gobs))) return rc;
    return PMC_OK;
}

// noise draw -> SPDE solve (from zero) -> [transfer] -> exp -> Darcy solve -> likelihood (and R = Q * likelihood)
static void emit_bayes_eval(Program &pg, Ctx *c, Rows &ar, Off mark, int level, Off rhs, Off k_ext, Off tbuf, Off Grow, Off Qrow,
                            Off out_row, bool with_q)
{
    SamplerLevel &S = c->s[level];
    DarcyLevel &D = c->d[level];
    const int Nk = S.out_size();
    {
        ar.top = mark;
        SolveWs ws;
        carve_solve(ar, S.sys, ws);
        emit_sampler_solve(pg, c, level, rhs, -1, ws, false);
        VecRef fsrc = vr(ws.x, S.sys.N, S.sys.Nf);
        if (S.hasT) {
            emit_spmm(pg, KC_TRANSFER, EP_AX, S.dT, VNULL, fsrc, vr(tbuf), VNULL, VNULL, nullptr, VNULL, 0, 0, -1, false, false,
                      (double)S.Ne + Nk);
            fsrc = vr(tbuf);
        }
        if (S.lognormal) { Op &o = pg.add(OP_MAP_EXP, KC_MISC, Nk, 2.0 * Nk); o.x = fsrc; o.y = vr(k_ext); }
        else emit_copy(pg, fsrc, vr(k_ext), Nk);
        emit_fill(pg, vr(k_ext, 0, Nk), 1, 1.0);
    }
    {
        ar.top = mark;
        SolveWs ws;
        carve_solve(ar, D.sys, ws);
        emit_darcy_solve(pg, c, level, k_ext, ws, with_q ? Qrow : (Off)-1, false);
        // G_i = g_i . p / sum(g_i)  (BayesianInverseProblem::ComputeG, /root/reference/src/BayesianInverseProblem.cpp:178-190)
        for (int i = 0; i < D.n_obs; ++i) {
            Op &o = pg.add(OP_DOT_FIXED, KC_MISC, D.Ne, D.Ne);
            o.fixed = D.d_gobs_func + (size_t)i * D.Ne;
            o.x = vr(ws.x, D.sys.N, D.sys.Nf);
            o.y = vr(Grow, 0, i);
        }
        Op &o = pg.add(OP_LIKELIHOOD, KC_MISC, D.n_obs, D.n_obs);
        o.x = vr(Grow);
        o.fixed = D.d_Gobs;
        o.ca = 1.0 / (2.0 * D.noise);  // exp((-1/(2 noise)) |G - G_obs|^2)  (:196-199)
        o.r = with_q ? vr(Qrow) : VNULL;
        o.y = vr(out_row);
    }
}

int pmc_bayes_level_batch(pmc_handle c, int level, int nlevels, int nsamples, uint64_t pos0, double *sums, double *rows,
                          int64_t *total_iters)
{
    if (!c) return PMC_ERR_ARG;
    if (!sums || nsamples < 0) return fail(c, PMC_ERR_ARG, "pmc_bayes_level_batch: bad arguments");
    if (nlevels < 1 || nlevels > c->nlevels || level < 0 || level >= nlevels)
        return fail(c, PMC_ERR_ARG, "level %d / nlevels %d out of range", level, nlevels);
    const bool coarsest = (level == nlevels - 1);
    int rc = check_level(c, level, true, true);
    if (rc) return rc;
    if (!coarsest) {
        if ((rc = check_level(c, level + 1, true, true))) return rc;
        if (!c->s[level].hasP) return fail(c, PMC_ERR_STATE, "sampler level %d has no prolongator", level);
    }
    for (int l = level; l <= (coarsest ? level : level + 1); ++l) {
        if (c->d[l].n_obs < 1) return fail(c, PMC_ERR_STATE, "no observations uploaded for level %d", l);
        if (c->s[l].out_size() != c->d[l].Ne) return fail(c, PMC_ERR_STATE, "sampler output / Darcy size mismatch on level %d", l);
    }
    if (!c->rng_ready) return fail(c, PMC_ERR_STATE, "pmc_rng_init has not been called");
    if (nsamples == 0) return PMC_OK;
    SamplerLevel &SF = c->s[level];
    DarcyLevel &DF = c->d[level];
    const int Ne = SF.Ne, Nec = coarsest ? 0 : c->s[level + 1].Ne;
    const int Nkmax = std::max(SF.out_size(), coarsest ? 0 : c->s[level + 1].out_size());
    const int mobs = std::max(DF.n_obs, coarsest ? 0 : c->d[level + 1].n_obs);
    // cost: every ComputeLikelihood / ComputeR adds the dofs of its solve (src/ML_BayesRatio_Manager.hpp:334-396)
    const double cost = 2.0 * DF.sys.N + (coarsest ? 0.0 : 2.0 * c->d[level + 1].sys.N);
    Rows ar;
    const Off rhs_f = ar.alloc(Ne), rhs_c = coarsest ? -1 : ar.alloc(Nec);
    const Off k_ext = ar.alloc(Nkmax + 1), tbuf = ar.alloc(Nkmax), Grow = ar.alloc(mobs), Qrow = ar.alloc(1);
    const Off o_z = ar.alloc(1), o_zc = ar.alloc(1), o_r = ar.alloc(1), o_rc = ar.alloc(1);
    const Off mark = ar.top;
    Program pg;
    pg.staging = c->staging;
    pg.defer_x = c->defer_x;
    pg.cheb3 = c->cheb3;
    pg.qoi_only = c->qoi_only;
    pg.fuse_coarse = c->fuse_coarse;
    pg.split_apply = c->split_apply;
    pg.chunked = c->stage_wide;
    std::vector<int> rng_ops;
    for (int draw = 0; draw < 2; ++draw) {  // draw 0: zxi -> Z (likelihood); draw 1: xi -> R = Q * likelihood
        rng_ops.push_back(pg.pc());
        {
            Op &o = pg.add(OP_RNG, KC_RNG, Ne, Ne);  // SamplePrior(ilevel, .): two vectors per realisation
            o.y = vr(rhs_f, Ne);
            o.fixed = SF.w_sqrt;
            o.ca = -SF.g;
            o.a0 = 2;
        }
        emit_bayes_eval(pg, c, ar, mark, level, rhs_f, k_ext, tbuf, Grow, Qrow, draw == 0 ? o_z : o_r, draw == 1);
        if (!coarsest) {
            // EvalPrior(ilevel+1, xi, .): 3-argument Eval, the right-hand side restricted with Ps^T
            emit_spmm(pg, KC_TRANSFER, EP_AX, SF.dPt, VNULL, vr(rhs_f, Ne), vr(rhs_c, Nec), VNULL, VNULL, nullptr, VNULL, 0, 0, -1,
                      false, false, (double)Ne + Nec);
            emit_bayes_eval(pg, c, ar, mark, level + 1, rhs_c, k_ext, tbuf, Grow, Qrow, draw == 0 ? o_zc : o_rc, draw == 1);
        }
    }
    const Off chunk = ar.peak;
    const size_t per_sample = (size_t)chunk * 8 / TW + 64 + 48;
    int B = 0;
    if ((rc = size_batch(c, per_sample, 40, nsamples, &B))) return rc;
    if ((rc = ensure_pinned(c, 32 + (rows ? (size_t)B * 5 : 0)))) return rc;
    unsigned long long it0 = c->iters_seen;   // replaced by the device counter read before the first launch (see level_batch)
    for (int s0 = 0; s0 < nsamples; s0 += B) {
        const int ns = std::min(B, nsamples - s0);
        const int ld = pad_ld(ns);
        double *out20 = (double *)c->arena.base + (size_t)(ld / TW) * (size_t)chunk;
        double *rows_d = rows ? out20 + 32 : nullptr;
        for (int draw = 0; draw < 2; ++draw) pg.ops[rng_ops[draw]].u0 = pos0 + (uint64_t)(2 * (uint64_t)s0 + draw) * (uint64_t)Ne;
        if (s0 == 0)
            CK(cudaMemcpyAsync(c->h_pinned + 21, &c->d_pstats->iters_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        if ((rc = run_program(c, pg, ns, chunk, DF.sys.N))) return rc;
        CK(cudaMemsetAsync(out20, 0, 20 * sizeof(double), c->stream));
        launch(c, PMC_K_MISC, (double)ns * 32.0, k_bayes_accumulate, dim3(1), dim3(256), ns, (const double *)c->arena.base, chunk,
               o_r, coarsest ? (Off)-1 : o_rc, o_z, coarsest ? (Off)-1 : o_zc, cost, out20, rows_d);
        CK(cudaMemcpyAsync(c->h_pinned, out20, 20 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(c->h_pinned + 20, &c->d_pstats->iters_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        if (rows) CK(cudaMemcpyAsync(c->h_pinned + 32, rows_d, (size_t)ns * 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if ((rc = finish(c))) return rc;
        if (s0 == 0) memcpy(&it0, c->h_pinned + 21, sizeof(unsigned long long));
        memcpy(&c->iters_seen, c->h_pinned + 20, sizeof(unsigned long long));
        for (int k = 0; k < 20; ++k) sums[k] += c->h_pinned[k];
        if (rows) memcpy(rows + 5 * (size_t)s0, c->h_pinned + 32, (size_t)ns * 5 * sizeof(double));
    }
    if (total_iters) *total_iters = (int64_t)(c->iters_seen - it0);
    return PMC_OK;
}

// ---- instrumentation ------------------------------------------------------------------------------
int pmc_reset_stats(pmc_handle c)
{
    if (!c) return PMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    resolve_events(c);
    memset(&c->stats, 0, sizeof c->stats);
    c->kernel_ms = 0.0;
    c->kernel_launches = 0;
    c->other_launches = 0;
    c->iters_seen = 0;
    CK(cudaMemsetAsync(c->d_pstats, 0, sizeof(ProgStats), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PMC_OK;
}

int pmc_kernel_stats(pmc_handle c, pmc_kernel_stats_t *out)
{
    if (!c || !out) return PMC_ERR_ARG;
    CK(cudaSetDevice(c->device));
    resolve_events(c);
    ProgStats ps;
    CK(cudaMemcpy(&ps, c->d_pstats, sizeof ps, cudaMemcpyDeviceToHost));
    pmc_kernel_stats_t st = c->stats;
    unsigned long long cyc = 0;
    for (int k = 0; k < KC_COUNT; ++k) cyc += ps.class_cycles[k];
    for (int k = 0; k < KC_COUNT; ++k) {
        st.class_cycle_share[k] = cyc ? (double)ps.class_cycles[k] / (double)cyc : 0.0;
        st.ms[k] = c->kernel_ms * st.class_cycle_share[k];
        st.algo_bytes[k] += ps.class_bytes[k];
        st.timed_launches[k] = (int64_t)ps.class_ops[k];
    }
    st.kernel_launches = c->kernel_launches;
    st.other_launches = c->other_launches;
    st.kernel_ms = c->kernel_ms;
    st.kernel_algo_bytes = ps.bytes;
    st.ops_executed = (int64_t)ps.ops_executed;
    st.minres_iterations = (int64_t)ps.iters_total;
    *out = st;
    return PMC_OK;
}

}  // extern "C"
