// program.cuh -- the tile-persistent solver kernel.
//
// EXECUTION MODEL.  The realisations of a batch are grouped into TILES of TW = 4 consecutive samples.  One CTA owns
// one tile for the whole life of a level batch: it runs a host-recorded PROGRAM (a flat list of operations: noise
// generation, restriction, per-solve value set-up, the preconditioned MINRES loop with its Chebyshev / V-cycle
// preconditioner, exp, QoI) from the first operation to the last, with only CTA-local barriers between operations
// and with the per-sample Krylov scalars, convergence flags and dot products living in shared memory.  There is one
// kernel launch per level batch, no grid-wide synchronisation and no host round trip inside a solve; a tile whose
// four realisations have converged leaves the Krylov loop on its own (no lock-step with the slowest sample of the
// batch).  Realisations never interact, so nothing is exchanged between CTAs.
//
// DATA LAYOUT.  Everything a tile touches lives in one contiguous chunk; inside it a batched vector with n rows is
// stored as X[row * TW + j], j = sample within the tile, i.e. n rows of 32 bytes.  Two adjacent threads share a row
// (one 128-bit access each), so a warp touches 16 consecutive rows = 512 contiguous bytes per access, the CTA
// streams each vector linearly, and the gathers of a sparse operator apply stay within a band of the tile's vector
// that lives in the SM's L1.  The sparse operators (CSR, weighted CSR) are shared by all tiles and stay in L2.  All
// traffic of the batched vectors goes to HBM once per operation (ncu: dram bytes = 1.03 x algorithmic bytes): the
// kernel is bound by HBM bandwidth and, per CTA, by memory latency -- hence CTA sizes chosen per level so that about
// 900 threads per SM are resident (72 registers per thread).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rng.cuh"

namespace pmc {

namespace cg = cooperative_groups;

constexpr int TW = 4;     // samples per tile (32-byte rows)
constexpr int LPR = 2;    // lanes per row: two adjacent threads share a row, 16 bytes (one double2) each
constexpr int PW = TW / LPR;
constexpr int SLICE = 32 / LPR;  // rows per sliced-ELL slice = rows one warp covers per pass
constexpr int MAXWARP = 16;  // largest CTA: 512 threads (the variants in use run at most 448)
constexpr int SMALLN = 64;   // OP_CHEB_SMALL: largest level whose iterates fit the staging buffers of the smallest CTA
constexpr int STW = 8;              // widest slice (entries per row) a staging buffer holds
constexpr int STCAP = STW * SLICE;  // entries per staging buffer
constexpr int NSTAGE = 2;           // staging buffers per warp

enum { EP_AX = 0, EP_RESID = 1, EP_CHEB = 2, EP_ADD = 3 };

enum {
    ST_BETA = 0, ST_IB, ST_IBPREV, ST_G0, ST_G1, ST_S0, ST_S1, ST_ETA, ST_GOAL, ST_ALPHA,
    ST_CQ, ST_CV1, ST_CV0, ST_CW0, ST_CW1, ST_CU, ST_CX, ST_CXP,
    ST_RZ, ST_CGA, ST_CGB,   // preconditioned CG: r.z, step length, direction coefficient
    ST_OM0, ST_OM1, ST_QACC, // MINRES for a linear functional of the solution only: obs . w (two generations) and obs . x
    ST_COUNT
};

enum OpKind {
    OP_SPMM = 0,      // y = epilogue(A x)       flags: ep | weighted | bdinv | dot
    OP_CHEB_FIRST,    // d = z = cb * dinv * r   flags: bdinv | dot
    OP_LINCOMB3,      // y = cq x + cv1 r + cv0 y  (per-sample coefficients); with F_DOT: fused mass-block Jacobi
                      //   d[row < a0] = cb * dinv * y, dots[slot] = y . d over those rows
    OP_SOL_UPDATE,    // y = cw0 y + cw1 r + cu x ; d += cx y   (y = w0, r = w1, x = u1, d = solution)
                      //   a0 = 1: the solution update is deferred (its coefficient is kept in ST_CXP)
                      //   a0 = 2: d += cxp r + cx y  (the deferred update of the previous iteration, then this one)
                      //   a0 = 3: d += cxp y only    (flush of a deferred update when the loop is left)
    OP_SETUP_SPMM,    // y = f(A g(x))           flags: absx | recip
    OP_FILL,          // y = ca
    OP_COPY,          // y = x
    OP_BROADCAST,     // y[row][j] = sample valid ? fixed[row] : 0
    OP_MAP_EXP,       // y = exp(x)
    OP_DOT_FIXED,     // y[0][j] = sum_row fixed[row] * x[row][j]   (one chunk row of per-sample results)
    OP_SC_INIT,       // Krylov state from dots[slot]
    OP_SC_ALPHA,
    OP_SC_BETA,
    OP_CHECK,         // if no sample of the tile is active: pc = a0
    OP_JUMP,          // pc = a0
    OP_STORE_ITERS,   // y[0][j] = iterations of sample j (as double) ; total += iterations
    OP_RNG,           // y[row][j] = (-g * N(mu,sigma)) * w_sqrt[row]  at stream position u0 + sample * a0 * n + row
    OP_CHEB_SMALL,    // whole Chebyshev iteration z = p(A) r on a level of at most SMALLN rows, iterates in shared memory
    OP_LIKELIHOOD,    // y[0][j] = exp(-sum_i (x[i][j] - fixed[i])^2 * ca) [* r[0][j]]   (n = number of observations)
    // Jacobi-preconditioned CG on the SPD form of the sampler system (emit_sampler_pcg in pmc_b200.cu)
    OP_CG_INIT,       // rz = dots[slot]; goal = max(rel sqrt(rz), abs); active = sqrt(rz) > goal
    OP_CG_ALPHA,      // a = active ? rz / dots[slot] : 0          (dots[slot] = p . H p)
    OP_CG_UPDATE,     // d += a x ; y -= a r ; dots[slot] = sum fixed[row] y[row]^2   (d = solution, x = p, y = residual, r = H p)
    OP_CG_BETA,       // b = active ? dots[slot] / rz : 0; rz = dots[slot]; iterations, convergence
    OP_CG_DIR,        // y = fixed[row] x + b y                     (y = p, x = residual)
    // Chebyshev semi-iteration on the SPD form of the sampler system (emit_sampler_cheb in pmc_b200.cu)
    OP_CHB_INIT,      // eta0 = sqrt(ca dots[slot]); goal = max(rel eta0, abs); active = valid && eta0 > goal; iterations = 0
    OP_CHB_CHECK,     // iterations += a0; active &= sqrt(dots[slot]) > goal && iterations < max_iter   (dots[slot] = r . D^-1 r)
    OP_STORE_Q,       // y[0][j] = obs . x accumulated by the functional-only MINRES (OP_SC_BETA with a1 = 1)
    OP_DOT_SPARSE,    // dots[3][j] = sum_i val[i] * x[col[i]][j]   (n = non-zero entries of a sparse fixed functional)
    OP_KIND_COUNT
};

enum {
    F_WEIGHTED = 1, F_BDINV = 2, F_DOT = 4, F_DOT_ACC = 8, F_DOT_WITH_R = 16, F_ABSX = 32, F_RECIP = 64,
    F_EP_SHIFT = 8,  // ep stored in bits 8..9
    F_STAGED = 1024,  // no slice of the operator is wider than STW: its entries are staged through shared memory
    // grid-group mode only (a tile owned by G CTAs of a cooperative launch):
    F_SOLO = 2048,       // small operation: executed by the group's first CTA alone
    F_LOCAL_SYNC = 4096, // nothing this operation writes is read by another CTA before the next group barrier
    F_CHUNKED = 16384,   // slices wider than STW: their entries are staged chunk by chunk (op_spmm_chunked)
    F_INLOOP = 8192,     // inside a Krylov loop: its bytes are credited for the realisations of the tile still iterating only
    F_THREE = 32768      // EP_CHEB in three-term form: the update d = z - z_prev is taken from the output buffer (which holds the
                         // previous iterate) instead of a separate vector: y = x + ca (x - y) + cb dinv (r - A x)
};

// kernel classes for the in-kernel time/byte accounting (same meaning as PMC_K_* in include/pmc_b200.h)
enum { KC_SADDLE = 0, KC_LANCZOS, KC_SOLUPD, KC_MASS, KC_SCHUR, KC_TRANSFER, KC_SETUP, KC_SCALAR, KC_RNG, KC_MISC, KC_COUNT };

// A batched operand: offset (in doubles) of the first row of the view inside the tile's CHUNK.  Everything a tile
// touches -- Krylov vectors, preconditioner scratch, right-hand sides, per-sample outputs -- lives in one contiguous
// chunk of `ProgParams::chunk` doubles at base + tile * chunk, so workspace reuse between the consecutive solves of a
// program is private to the tile (tiles progress independently of one another).  off < 0: no operand.
struct VecRef {
    long long off;
};

struct Op {
    int kind, flags, n, slot;
    int a0, a1, kclass, pad;
    // OP_SPMM: sliced-ELL arrays (slice = 16 rows = one warp pass; entry k of slice s, row r: index (k * 16 + r) with
    // k in [rowptr[s], rowptr[s+1])); weighted operators carry a weight index per entry (their sample-independent
    // entries point at a weight row that holds the constant 1).
    // OP_SETUP_SPMM: plain CSR arrays.
    const int *rowptr;
    const int *col;
    const double *val;
    const unsigned char *pk;
    const double *fixed;  // fixed (sample-independent) vector: 1/diag, obs, rhs, w_sqrt
    VecRef x, y, r, d, w, v;
    double ca, cb, bytes;  // bytes: algorithmic bytes this op moves per tile
    unsigned long long u0;
};

struct ProgStats {
    unsigned long long class_cycles[KC_COUNT];  // summed over CTAs
    double class_bytes[KC_COUNT];               // algorithmic bytes per class, summed over tiles
    unsigned long long class_ops[KC_COUNT];
    unsigned long long ops_executed;
    unsigned long long iters_total;
    double bytes;                               // algorithmic bytes of the executed ops, summed over tiles
    unsigned long long cta_cycles;              // summed over CTAs
};

struct ProgParams {
    const Op *ops;
    int nops, ntiles, nsamples, max_iter;
    double rel, abs_, mu, sigma;
    const RngTables *tab;
    ProgStats *stats;
    double *base;      // tile chunks
    long long chunk;   // doubles per tile chunk
    int group;                 // grid-group mode (CS = 0): CTAs per tile
    unsigned int *grp_bar;     //   [ntiles][2] barrier words, zeroed before the launch
    double *grp_part;          //   [ntiles][group][TW]
    unsigned long long *op_cycles;  // diagnostic (PMC_OP_PROFILE): per-operation {cycles, executions} summed over CTAs, or null
};

struct Smem {
    double st[ST_COUNT][TW];
    double dots[4][TW];
    double red[MAXWARP][TW];
    double part[TW];   // this CTA's share of a dot product (read by the other CTAs of the cluster through DSMEM)
    int active[TW];
    int iters[TW];
    unsigned long long bars[MAXWARP][NSTAGE];  // per-warp mbarriers of the operator staging buffers
    int grp_size, grp_rank;                    // grid-group mode: CTAs per tile and this CTA's rank among them
    unsigned int *grp_bar;                     //   {count, generation} of the tile's barrier (global memory)
    double *grp_part;                          //   [grp_size][TW] dot-product shares of the tile (global memory)
    unsigned long long cyc[KC_COUNT];
    double cbytes[KC_COUNT];
    unsigned int cops[KC_COUNT];
};

typedef double2 D2;
__device__ __forceinline__ D2 ld2c(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, D2 v) { *reinterpret_cast<double2 *>(p) = v; }
__device__ __forceinline__ double *tp(const VecRef &v, double *chunk) { return v.off >= 0 ? chunk + v.off : nullptr; }
__device__ __forceinline__ double safe_inv(double x) { return x != 0.0 ? 1.0 / x : 0.0; }

// CLUSTER SPLIT.  When a batch has too few tiles to fill the machine (a fine level with a handful of realisations, or
// very large levels where memory limits the number of tiles), a tile is owned by a thread-block CLUSTER of CS CTAs:
// every operation's rows are split into CS contiguous ranges (multiples of a slice), the barrier between operations
// becomes a cluster barrier, and dot products are combined through distributed shared memory in rank order, so every
// CTA of the cluster holds the same scalars and takes the same branches.  CS = 1 compiles to plain CTA barriers.
//
// GRID GROUPS (CS = 0).  A cluster is at most 8 CTAs, so a batch of one or two tiles of a very large level (SPE10's
// 4.5 M-row fine level holds 2 GB per tile) would still leave most of the machine idle.  In this mode a tile is owned
// by a GROUP of G CTAs of a cooperative launch (all co-resident): rows are split G ways, the barrier between
// operations is a sense-reversing barrier on a word in global memory (release/acquire fences at GPU scope, which also
// invalidate the SM's L1), dot-product shares travel through global memory and are added in rank order by every CTA.
// Operations too small to be worth a group barrier (the coarse levels of the V-cycle) are executed by the group's
// first CTA alone (F_SOLO) and chained with CTA-local barriers (F_LOCAL_SYNC); scalar operations only touch shared
// memory and never need the group barrier.
template <int CS>
__device__ __forceinline__ int cluster_rank(const Smem &sm)
{
    if (CS == 1) return 0;
    if (CS == 0) return sm.grp_rank;
    return (int)cg::this_cluster().block_rank();
}
template <int CS>
__device__ __forceinline__ void my_rows(const Op &o, const Smem &sm, int &r0, int &r1)
{
    const int n = o.n;
    if (CS == 1) { r0 = 0; r1 = n; return; }
    if (CS == 0 && (o.flags & F_SOLO)) { r0 = 0; r1 = sm.grp_rank == 0 ? n : 0; return; }
    const int parts = CS == 0 ? sm.grp_size : CS;
    const int per = (((n + parts - 1) / parts) + SLICE - 1) & ~(SLICE - 1);
    r0 = min(n, cluster_rank<CS>(sm) * per);
    r1 = min(n, r0 + per);
}
// barrier of the tile's grid group: every thread of every CTA of the group calls it
__device__ __forceinline__ void group_barrier(const Smem &sm)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int *gen_p = sm.grp_bar + 1;
        const unsigned int gen = *gen_p;  // cannot advance before this CTA has arrived
        __threadfence();
        if (atomicAdd(sm.grp_bar, 1u) == (unsigned int)sm.grp_size - 1u) {
            *(volatile unsigned int *)sm.grp_bar = 0u;
            __threadfence();
            atomicAdd(sm.grp_bar + 1, 1u);
        } else {
            while (*gen_p == gen) __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}
template <int CS>
__device__ __forceinline__ void op_barrier(const Smem &sm, int flags)
{
    if (CS == 1) __syncthreads();
    else if (CS == 0) {
        if (flags & F_LOCAL_SYNC) __syncthreads();
        else group_barrier(sm);
    } else cg::this_cluster().sync();
}

// Deterministic reduction.  Every thread holds the partial sums of its PW samples (lane parity selects which
// samples); xor-butterfly over lanes of equal parity, the warp results are added in warp order, and with a cluster the
// CTA results in rank order.  Result lands in sm.dots[slot] (= or +=) of every CTA.
template <int NTt, int CS>
__device__ __forceinline__ void block_dot(D2 acc, Smem &sm, int slot, bool accumulate)
{
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < LPR) {
        sm.red[warp][lane * PW + 0] = acc.x;
        sm.red[warp][lane * PW + 1] = acc.y;
    }
    __syncthreads();
    if (CS == 1) {
        if (threadIdx.x < TW) {
            double s = accumulate ? sm.dots[slot][threadIdx.x] : 0.0;
#pragma unroll
            for (int w = 0; w < NTt / 32; ++w) s += sm.red[w][threadIdx.x];
            sm.dots[slot][threadIdx.x] = s;
        }
    } else if (CS == 0) {
        if (threadIdx.x < TW) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NTt / 32; ++w) s += sm.red[w][threadIdx.x];
            sm.grp_part[sm.grp_rank * TW + threadIdx.x] = s;
        }
        group_barrier(sm);
        if (threadIdx.x < TW) {
            double s = accumulate ? sm.dots[slot][threadIdx.x] : 0.0;
            for (int r = 0; r < sm.grp_size; ++r) s += __ldcg(sm.grp_part + r * TW + threadIdx.x);
            sm.dots[slot][threadIdx.x] = s;
        }
        // the group barrier that ends the operation keeps the shares alive until every CTA has read them
    } else {
        if (threadIdx.x < TW) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NTt / 32; ++w) s += sm.red[w][threadIdx.x];
            sm.part[threadIdx.x] = s;
        }
        cg::cluster_group cl = cg::this_cluster();
        cl.sync();
        if (threadIdx.x < TW) {
            double s = accumulate ? sm.dots[slot][threadIdx.x] : 0.0;
            for (int r = 0; r < CS; ++r) s += *cl.map_shared_rank(&sm.part[threadIdx.x], r);
            sm.dots[slot][threadIdx.x] = s;
        }
        // the barrier that ends the operation keeps `part` alive until every CTA has read it
    }
}

// ---- operator staging ------------------------------------------------------------------------------------
// A sparse apply walks the operator slice by slice (one slice = the 16 rows a warp covers per pass).  Read straight
// from L2, every pass is a chain of dependent long-latency loads: slice offsets -> entries (value, column, weight
// index) -> gathers of the batched vector -> fma.  The entries of a slice are contiguous in the packed sliced-ELL
// array, so each warp keeps NSTAGE passes of entries in flight with TMA bulk copies (cp.async.bulk, completion on a
// per-warp mbarrier) into its own shared-memory buffers: the chain seen by the gathers shrinks to shared-memory reads.
// No CTA-wide synchronisation is involved; warps stay decoupled.
struct WarpStage {
    unsigned char bytes[STCAP * 16];  // [val: w*16 doubles][col: w*16 ints][widx: w*16 ints], w <= STW
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// per-warp staging state (registers; the same buffers and barriers serve every staged operation of the program)
struct StageCtx {
    WarpStage *buf;   // this warp's NSTAGE buffers
    uint32_t bar;     // shared address of this warp's NSTAGE mbarriers
    uint32_t phase;   // bit s: parity the next wait on buffer s expects
};

template <int ES>
__device__ __forceinline__ void stage_issue(const unsigned char *pk, WarpStage *dst, uint32_t bar, int k0, int k1)
{
    if ((threadIdx.x & 31) == 0) {
        const uint32_t bytes = (uint32_t)(k1 - k0) * (SLICE * ES);
        mbar_arrive_tx(bar, bytes);
        if (bytes) bulk_g2s(smem_u32(dst), pk + (size_t)k0 * (SLICE * ES), bytes, bar);
    }
}

// One row's entries, compile-time width: sum_k val_k [V[widx_k]] x[col_k] in entry order.
template <bool WEIGHTED, int WF>
__device__ __forceinline__ D2 gather_fixed(const double *__restrict__ eval, const int *__restrict__ ecol, const int *__restrict__ ewid,
                                           const double *__restrict__ x, const double *__restrict__ V)
{
    D2 s = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < WF; ++k) {
        const double c = eval[k * SLICE];
        const D2 xv = ld2c(x + (size_t)ecol[k * SLICE] * TW);
        if (WEIGHTED) {
            const D2 wv = ld2c(V + (size_t)ewid[k * SLICE] * TW);
            s.x = fma(c * wv.x, xv.x, s.x);
            s.y = fma(c * wv.y, xv.y, s.y);
        } else {
            s.x = fma(c, xv.x, s.x);
            s.y = fma(c, xv.y, s.y);
        }
    }
    return s;
}

// ---- sparse operator apply with fused epilogue ------------------------------------------------------------
//   plain    : row i of slice s: entries k in [off[s], off[s+1])     sum += val * x[col]
//   weighted :                                                        sum += val * V[widx] * x[col]
//   V = per-sample weights: the permeability k_e for M(k) = sum_e k_e R_e^T M_e R_e (the element reassembly of
//   DarcySolver::assemble, /root/reference/src/DarcySolver.cpp:479, never materialised) or the Schur values; entries
//   that do not depend on the sample (B, B^T, identity rows) point at a weight row holding the constant 1.
//   Packed sliced ELL: slice s occupies bytes [off[s], off[s+1]) * 16 * ES of `pk` (ES = 16 weighted, 12 plain) as
//   [val: w*16 doubles][col: w*16 ints][widx: w*16 ints], entry k of row r at index k * 16 + r.
template <int NTt, int CS, int EP, bool WEIGHTED, bool BDINV, bool DOT, bool STAGED, bool THREE = false>
__device__ __forceinline__ void op_spmm(const Op &o, double *chunk, Smem &sm, StageCtx &sc)
{
    constexpr int ES = WEIGHTED ? 16 : 12;
    constexpr int NW = NTt / 32;
    const int sub = (threadIdx.x % LPR) * PW;  // first sample of this thread inside the row
    const double *__restrict__ x = tp(o.x, chunk) + sub;
    double *__restrict__ y = tp(o.y, chunk) + sub;
    const double *__restrict__ r = (EP == EP_RESID || EP == EP_CHEB) ? tp(o.r, chunk) + sub : nullptr;
    // EP_AX with a fused dot: the partner of row i is x[i], or -- when the rows of this operation are a block of a larger
    // operator whose columns start elsewhere -- the vector given in o.r
    const double *__restrict__ xd = (EP == EP_AX && DOT && o.r.off >= 0) ? tp(o.r, chunk) + sub : x;
    double *__restrict__ d = (EP == EP_CHEB && !THREE) ? tp(o.d, chunk) + sub : nullptr;
    const double *__restrict__ V = WEIGHTED ? tp(o.v, chunk) + sub : nullptr;
    const double *__restrict__ dinvb = (EP == EP_CHEB && BDINV) ? tp(o.w, chunk) + sub : nullptr;
    const int *__restrict__ off = o.rowptr;
    const unsigned char *__restrict__ pk = o.pk;
    const double ca = o.ca, cb = o.cb;
    const bool dot_r = (o.flags & F_DOT_WITH_R) != 0;
    D2 acc = make_double2(0.0, 0.0);
    int r0, r1;
    my_rows<CS>(o, sm, r0, r1);
    // whole warps stay converged through the slice loops; an empty range (r0 = r1 = n need not be slice-aligned) has no slices
    const int sl_end = r1 > r0 ? (r1 + SLICE - 1) / SLICE : 0;
    const int rs = (threadIdx.x & 31) / LPR;
    int sl = r0 / SLICE + (threadIdx.x >> 5);
    int wcur = 0, wnext = 0, st = 0;
    if (STAGED) {
#pragma unroll
        for (int j = 0; j < NSTAGE; ++j) {
            const int slj = sl + j * NW;
            int k0 = 0, k1 = 0;
            if (slj < sl_end) {
                k0 = __ldg(off + slj);
                k1 = __ldg(off + slj + 1);
                stage_issue<ES>(pk, sc.buf + j, sc.bar + 8 * j, k0, k1);
            }
            if (j == 0) wcur = k1 - k0;
            else wnext = k1 - k0;
        }
    }
#ifdef PMC_SPMM_SYNC   // experiment: keep the warps of a CTA within a window of passes (L1 reuse between adjacent slices)
    const int maxp = (sl_end - r0 / SLICE + NW - 1) / NW;   // passes of the busiest warp: the same for every thread
    int pass = 0;
#endif
    for (; sl < sl_end; sl += NW) {
        const int row = sl * SLICE + rs;
        const bool live = row < r1;
        const size_t ro = (size_t)row * TW;
        // slice offsets of the pass that will reuse this buffer, and the streamed operands of the epilogue: issued
        // before the gathers so that their latency overlaps them
        const int sl2 = sl + NSTAGE * NW;
        int n0 = 0, n1 = 0;
        if (STAGED && sl2 < sl_end) {
            n0 = __ldg(off + sl2);
            n1 = __ldg(off + sl2 + 1);
        }
        D2 rv = make_double2(0.0, 0.0), di = rv, dv = rv, xr = rv, yv = rv;
#ifdef PMC_HOIST_EPILOGUE   // measured: loading the epilogue operands after the gathers is 1 % faster (registers)
        if (live) {
#else
        if (false) {
#endif
            if (EP == EP_RESID || EP == EP_CHEB) rv = ld2c(r + ro);
            if (EP == EP_ADD) yv = ld2c(y + ro);
            if (EP == EP_CHEB) {
                if (BDINV) di = ld2c(dinvb + ro);
                else { const double t = __ldg(o.fixed + row); di = make_double2(t, t); }
                if (ca != 0.0) dv = THREE ? ld2c(y + ro) : ld2c(d + ro);
            }
            if (EP == EP_CHEB) xr = ld2c(x + ro);
            else if (DOT && !dot_r) xr = ld2c(xd + ro);
        }
        const unsigned char *base;
        int w;
        if (STAGED) {
            mbar_wait(sc.bar + 8 * st, (sc.phase >> st) & 1u);
            sc.phase ^= 1u << st;
            base = sc.buf[st].bytes;
            w = wcur;
        } else {
            const int k0 = __ldg(off + sl);
            w = __ldg(off + sl + 1) - k0;
            base = pk + (size_t)k0 * (SLICE * ES);
        }
        const double *__restrict__ eval = reinterpret_cast<const double *>(base) + rs;
        const int *__restrict__ ecol = reinterpret_cast<const int *>(base + (size_t)w * (SLICE * 8)) + rs;
        const int *__restrict__ ewid = ecol + w * SLICE;
        D2 s;
#ifndef PMC_NO_FIXED_WIDTHS
        // the widths of the structured hot operators get fully unrolled bodies
        if (w == 7) s = gather_fixed<WEIGHTED, 7>(eval, ecol, ewid, x, V);
        else if (w == 6) s = gather_fixed<WEIGHTED, 6>(eval, ecol, ewid, x, V);
        else if (w == 5) s = gather_fixed<WEIGHTED, 5>(eval, ecol, ewid, x, V);
        else
#endif
        {
            s = make_double2(0.0, 0.0);
#pragma unroll 4
            for (int k = 0; k < w; ++k) {
                const double c = eval[k * SLICE];
                const D2 xv = ld2c(x + (size_t)ecol[k * SLICE] * TW);
                if (WEIGHTED) {
                    const D2 wv = ld2c(V + (size_t)ewid[k * SLICE] * TW);
                    s.x = fma(c * wv.x, xv.x, s.x);
                    s.y = fma(c * wv.y, xv.y, s.y);
                } else {
                    s.x = fma(c, xv.x, s.x);
                    s.y = fma(c, xv.y, s.y);
                }
            }
        }
        if (live) {
#ifndef PMC_HOIST_EPILOGUE
            if (EP == EP_RESID || EP == EP_CHEB) rv = ld2c(r + ro);
            if (EP == EP_ADD) yv = ld2c(y + ro);
            if (EP == EP_CHEB) {
                if (BDINV) di = ld2c(dinvb + ro);
                else { const double t = __ldg(o.fixed + row); di = make_double2(t, t); }
                if (ca != 0.0) dv = THREE ? ld2c(y + ro) : ld2c(d + ro);
            }
            if (EP == EP_CHEB) xr = ld2c(x + ro);
            else if (DOT && !dot_r) xr = ld2c(xd + ro);
#endif
            D2 out;
            if (EP == EP_AX) {
                out = s;
            } else if (EP == EP_RESID) {
                out = make_double2(rv.x - s.x, rv.y - s.y);
            } else if (EP == EP_ADD) {
                out = make_double2(fma(ca, s.x, yv.x), fma(ca, s.y, yv.y));
            } else {  // EP_CHEB: d = ca d + cb dinv (r - A z);  z_out = z + d
                D2 dn = make_double2(cb * di.x * (rv.x - s.x), cb * di.y * (rv.y - s.y));
                if (ca != 0.0) {
                    if (THREE) dv = make_double2(xr.x - dv.x, xr.y - dv.y);
                    dn.x = fma(ca, dv.x, dn.x);
                    dn.y = fma(ca, dv.y, dn.y);
                }
                if (!THREE) st2(d + ro, dn);
                out = make_double2(xr.x + dn.x, xr.y + dn.y);
                if (DOT && dot_r) {
                    acc.x = fma(out.x, rv.x, acc.x);
                    acc.y = fma(out.y, rv.y, acc.y);
                }
            }
            st2(y + ro, out);
            if (DOT && !dot_r) {
                acc.x = fma(out.x, xr.x, acc.x);
                acc.y = fma(out.y, xr.y, acc.y);
            }
        }
        if (STAGED) {
            __syncwarp();  // every lane is done with buffer st before it is refilled
            if (sl2 < sl_end) stage_issue<ES>(pk, sc.buf + st, sc.bar + 8 * st, n0, n1);
            wcur = wnext;
            wnext = n1 - n0;
            st ^= 1;
        }
#ifdef PMC_SPMM_SYNC
        if ((++pass % PMC_SPMM_SYNC) == 0) __syncthreads();
#endif
    }
#ifdef PMC_SPMM_SYNC
    while (pass < maxp)
        if ((++pass % PMC_SPMM_SYNC) == 0) __syncthreads();
#endif
    if (DOT) block_dot<NTt, CS>(acc, sm, o.slot, (o.flags & F_DOT_ACC) != 0);
}

// ---- sparse apply for operators with slices wider than the staging buffers (unstructured agglomerates, smoothed
// aggregation: rows of 10-40 entries) -------------------------------------------------------------------------------
// Same arithmetic as op_spmm in the same order (bitwise the same results as reading the entries from L2), but a wide
// slice is staged CHUNK by chunk of STW entries: the pipeline unit is (slice, chunk) instead of slice, each unit is three
// TMA bulk copies (values, columns, weight indices of the chunk are not adjacent in the packed slice), the row sum is
// carried across the chunks of a slice and the epilogue runs after the last one.
template <int NTt, int CS, int EP, bool WEIGHTED, bool BDINV, bool DOT, bool THREE = false>
__device__ __forceinline__ void op_spmm_chunked(const Op &o, double *chunk, Smem &sm, StageCtx &sc)
{
    constexpr int ES = WEIGHTED ? 16 : 12;
    constexpr int NW = NTt / 32;
    const int sub = (threadIdx.x % LPR) * PW;
    const double *__restrict__ x = tp(o.x, chunk) + sub;
    double *__restrict__ y = tp(o.y, chunk) + sub;
    const double *__restrict__ r = (EP == EP_RESID || EP == EP_CHEB) ? tp(o.r, chunk) + sub : nullptr;
    const double *__restrict__ xd = (EP == EP_AX && DOT && o.r.off >= 0) ? tp(o.r, chunk) + sub : x;
    double *__restrict__ d = (EP == EP_CHEB && !THREE) ? tp(o.d, chunk) + sub : nullptr;
    const double *__restrict__ V = WEIGHTED ? tp(o.v, chunk) + sub : nullptr;
    const double *__restrict__ dinvb = (EP == EP_CHEB && BDINV) ? tp(o.w, chunk) + sub : nullptr;
    const int *__restrict__ off = o.rowptr;
    const unsigned char *__restrict__ pk = o.pk;
    const double ca = o.ca, cb = o.cb;
    const bool dot_r = (o.flags & F_DOT_WITH_R) != 0;
    D2 acc = make_double2(0.0, 0.0);
    int r0, r1;
    my_rows<CS>(o, sm, r0, r1);
    const int sl_end = r1 > r0 ? (r1 + SLICE - 1) / SLICE : 0;
    const int rs = (threadIdx.x & 31) / LPR;
    const bool lane0 = (threadIdx.x & 31) == 0;
    // a unit: chunk c of slice sl (w entries from packed entry k0); sl >= sl_end: none
    int u_sl[3], u_c[3], u_w[3], u_k0[3];
    auto first_unit = [&](int j, int sl) {
        u_sl[j] = sl; u_c[j] = 0; u_w[j] = 0; u_k0[j] = 0;
        if (sl < sl_end) {
            u_k0[j] = __ldg(off + sl);
            u_w[j] = __ldg(off + sl + 1) - u_k0[j];
        }
    };
    auto next_unit = [&](int j, int i) {   // unit j := the one after unit i
        if ((u_c[i] + 1) * STW < u_w[i]) { u_sl[j] = u_sl[i]; u_c[j] = u_c[i] + 1; u_w[j] = u_w[i]; u_k0[j] = u_k0[i]; }
        else first_unit(j, u_sl[i] + NW);
    };
    auto issue = [&](int j, int buf) {
        if (!lane0 || u_sl[j] >= sl_end) return;
        const int kk = min(STW, u_w[j] - u_c[j] * STW);
        const uint32_t bar = sc.bar + 8 * buf;
        const uint32_t dst = smem_u32(sc.buf + buf);
        mbar_arrive_tx(bar, (uint32_t)max(kk, 0) * (SLICE * ES));
        if (kk <= 0) return;
        const unsigned char *base = pk + (size_t)u_k0[j] * (SLICE * ES);
        const size_t w = (size_t)u_w[j], c0 = (size_t)u_c[j] * STW;
        bulk_g2s(dst, base + c0 * (SLICE * 8), (uint32_t)kk * (SLICE * 8), bar);
        bulk_g2s(dst + kk * (SLICE * 8), base + w * (SLICE * 8) + c0 * (SLICE * 4), (uint32_t)kk * (SLICE * 4), bar);
        if (WEIGHTED)
            bulk_g2s(dst + kk * (SLICE * 12), base + w * (SLICE * 12) + c0 * (SLICE * 4), (uint32_t)kk * (SLICE * 4), bar);
    };
    first_unit(0, r0 / SLICE + (threadIdx.x >> 5));
    next_unit(1, 0);
    issue(0, 0);
    issue(1, 1);
    int st = 0;
    D2 s = make_double2(0.0, 0.0);
    while (u_sl[0] < sl_end) {
        next_unit(2, 1);
        mbar_wait(sc.bar + 8 * st, (sc.phase >> st) & 1u);
        sc.phase ^= 1u << st;
        const int kk = max(0, min(STW, u_w[0] - u_c[0] * STW));
        const unsigned char *base = sc.buf[st].bytes;
        const double *__restrict__ eval = reinterpret_cast<const double *>(base) + rs;
        const int *__restrict__ ecol = reinterpret_cast<const int *>(base + (size_t)kk * (SLICE * 8)) + rs;
        const int *__restrict__ ewid = ecol + kk * SLICE;
#pragma unroll 4
        for (int k = 0; k < kk; ++k) {
            const double c = eval[k * SLICE];
            const D2 xv = ld2c(x + (size_t)ecol[k * SLICE] * TW);
            if (WEIGHTED) {
                const D2 wv = ld2c(V + (size_t)ewid[k * SLICE] * TW);
                s.x = fma(c * wv.x, xv.x, s.x);
                s.y = fma(c * wv.y, xv.y, s.y);
            } else {
                s.x = fma(c, xv.x, s.x);
                s.y = fma(c, xv.y, s.y);
            }
        }
        __syncwarp();   // every lane is done with buffer st before it is refilled
        issue(2, st);
        if ((u_c[0] + 1) * STW >= u_w[0]) {   // last chunk of the slice: epilogue
            const int row = u_sl[0] * SLICE + rs;
            if (row < r1) {
                const size_t ro = (size_t)row * TW;
                D2 rv = make_double2(0.0, 0.0), di = rv, dv = rv, xr = rv, yv = rv;
                if (EP == EP_RESID || EP == EP_CHEB) rv = ld2c(r + ro);
                if (EP == EP_ADD) yv = ld2c(y + ro);
                if (EP == EP_CHEB) {
                    if (BDINV) di = ld2c(dinvb + ro);
                    else { const double t = __ldg(o.fixed + row); di = make_double2(t, t); }
                    if (ca != 0.0) dv = THREE ? ld2c(y + ro) : ld2c(d + ro);
                    xr = ld2c(x + ro);
                } else if (DOT && !dot_r) xr = ld2c(xd + ro);
                D2 out;
                if (EP == EP_AX) out = s;
                else if (EP == EP_RESID) out = make_double2(rv.x - s.x, rv.y - s.y);
                else if (EP == EP_ADD) out = make_double2(fma(ca, s.x, yv.x), fma(ca, s.y, yv.y));
                else {
                    D2 dn = make_double2(cb * di.x * (rv.x - s.x), cb * di.y * (rv.y - s.y));
                    if (ca != 0.0) {
                        if (THREE) dv = make_double2(xr.x - dv.x, xr.y - dv.y);
                        dn.x = fma(ca, dv.x, dn.x);
                        dn.y = fma(ca, dv.y, dn.y);
                    }
                    if (!THREE) st2(d + ro, dn);
                    out = make_double2(xr.x + dn.x, xr.y + dn.y);
                    if (DOT && dot_r) {
                        acc.x = fma(out.x, rv.x, acc.x);
                        acc.y = fma(out.y, rv.y, acc.y);
                    }
                }
                st2(y + ro, out);
                if (DOT && !dot_r) {
                    acc.x = fma(out.x, xr.x, acc.x);
                    acc.y = fma(out.y, xr.y, acc.y);
                }
            }
            s = make_double2(0.0, 0.0);
        }
        u_sl[0] = u_sl[1]; u_c[0] = u_c[1]; u_w[0] = u_w[1]; u_k0[0] = u_k0[1];
        u_sl[1] = u_sl[2]; u_c[1] = u_c[2]; u_w[1] = u_w[2]; u_k0[1] = u_k0[2];
        st ^= 1;
    }
    if (DOT) block_dot<NTt, CS>(acc, sm, o.slot, (o.flags & F_DOT_ACC) != 0);
}

template <int NTt, int CS, bool BDINV>
__device__ __forceinline__ void op_cheb_first(const Op &o, double *chunk, Smem &sm)
{
    const int sub = (threadIdx.x % LPR) * PW;
    int r0, r1;
    my_rows<CS>(o, sm, r0, r1);
    const double *__restrict__ r = tp(o.r, chunk) + sub;
    double *__restrict__ d = tp(o.d, chunk) + sub;
    double *__restrict__ z = tp(o.y, chunk) + sub;
    const double *__restrict__ dinvb = BDINV ? tp(o.w, chunk) + sub : nullptr;
    const bool dot = (o.flags & F_DOT) != 0;
    D2 acc = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
        const size_t ro = (size_t)row * TW;
        const D2 rv = ld2c(r + ro);
        D2 di;
        if (BDINV) di = ld2c(dinvb + ro);
        else { const double t = __ldg(o.fixed + row); di = make_double2(t, t); }
        const D2 dn = make_double2(o.cb * di.x * rv.x, o.cb * di.y * rv.y);
        acc.x = fma(dn.x, rv.x, acc.x);
        acc.y = fma(dn.y, rv.y, acc.y);
        st2(d + ro, dn);
        st2(z + ro, dn);
    }
    if (dot) block_dot<NTt, CS>(acc, sm, o.slot, (o.flags & F_DOT_ACC) != 0);
}

// Lanczos update v0 = cq q + cv1 v1 + cv0 v0, optionally fused with the Jacobi preconditioner of the RT mass block
// (rows < a0): z = cb * dinv * v0 written to o.d, and the partial dot v0 . z of those rows.
template <int NTt, int CS, bool JACOBI, bool BDINV>
__device__ __forceinline__ void op_lincomb3(const Op &o, double *chunk, Smem &sm)
{
    const int sub = (threadIdx.x % LPR) * PW;
    int r0, r1;
    my_rows<CS>(o, sm, r0, r1);
    const double *__restrict__ q = tp(o.x, chunk) + sub;
    const double *__restrict__ v1 = tp(o.r, chunk) + sub;
    double *__restrict__ v0 = tp(o.y, chunk) + sub;
    double *__restrict__ z = JACOBI ? tp(o.d, chunk) + sub : nullptr;
    const double *__restrict__ dinvb = (JACOBI && BDINV) ? tp(o.w, chunk) + sub : nullptr;
    const D2 a = make_double2(sm.st[ST_CQ][sub], sm.st[ST_CQ][sub + 1]);
    const D2 b = make_double2(sm.st[ST_CV1][sub], sm.st[ST_CV1][sub + 1]);
    const D2 c = make_double2(sm.st[ST_CV0][sub], sm.st[ST_CV0][sub + 1]);
    const bool anyc = c.x != 0.0 || c.y != 0.0;
    const int nj = o.a0;
    const double cb = o.cb;
    D2 acc = make_double2(0.0, 0.0);
    // rows [r0, rj): with the fused Jacobi block; rows [rj, r1): the plain update.  Two branch-free loops (the loads of a
    // conditional body cannot be hoisted over the stores of the previous row); a thread keeps the rows it had in the
    // single loop, so the partial sums of the dot are added in the same order.
    const int rj = JACOBI ? min(r1, max(r0, nj)) : r0;
    int row = r0 + threadIdx.x / LPR;
    if (JACOBI) {
#pragma unroll 4
        for (; row < rj; row += NTt / LPR) {
            const size_t ro = (size_t)row * TW;
            const D2 qv = ld2c(q + ro), vv = ld2c(v1 + ro);
            D2 cv = make_double2(0.0, 0.0);
            if (anyc) cv = ld2c(v0 + ro);
            D2 di;
            if (BDINV) di = ld2c(dinvb + ro);
            else { const double t = __ldg(o.fixed + row); di = make_double2(t, t); }
            const D2 vn = make_double2(fma(a.x, qv.x, fma(b.x, vv.x, c.x * cv.x)), fma(a.y, qv.y, fma(b.y, vv.y, c.y * cv.y)));
            st2(v0 + ro, vn);
            const D2 zn = make_double2(cb * di.x * vn.x, cb * di.y * vn.y);
            st2(z + ro, zn);
            acc.x = fma(zn.x, vn.x, acc.x);
            acc.y = fma(zn.y, vn.y, acc.y);
        }
    }
#pragma unroll 4
    for (; row < r1; row += NTt / LPR) {
        const size_t ro = (size_t)row * TW;
        const D2 qv = ld2c(q + ro), vv = ld2c(v1 + ro);
        D2 cv = make_double2(0.0, 0.0);
        if (anyc) cv = ld2c(v0 + ro);
        const D2 vn = make_double2(fma(a.x, qv.x, fma(b.x, vv.x, c.x * cv.x)), fma(a.y, qv.y, fma(b.y, vv.y, c.y * cv.y)));
        st2(v0 + ro, vn);
    }
    if (JACOBI) block_dot<NTt, CS>(acc, sm, o.slot, false);
}

template <int NTt, int CS>
__device__ __forceinline__ void op_sol_update(const Op &o, double *chunk, Smem &sm)
{
    const int sub = (threadIdx.x % LPR) * PW;
    int r0, r1;
    my_rows<CS>(o, sm, r0, r1);
    double *__restrict__ w0 = tp(o.y, chunk) + sub;
    const int mode = o.a0;
    const D2 ep = make_double2(sm.st[ST_CXP][sub], sm.st[ST_CXP][sub + 1]);
    if (mode == 3) {  // x += cxp w0
        double *__restrict__ xs = tp(o.d, chunk) + sub;
        if (ep.x == 0.0 && ep.y == 0.0) return;
#pragma unroll 4
        for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
            const size_t ro = (size_t)row * TW;
            const D2 wv = ld2c(w0 + ro);
            D2 xv = ld2c(xs + ro);
            if (ep.x != 0.0) xv.x = fma(ep.x, wv.x, xv.x);
            if (ep.y != 0.0) xv.y = fma(ep.y, wv.y, xv.y);
            st2(xs + ro, xv);
        }
        return;
    }
    const double *__restrict__ w1 = tp(o.r, chunk) + sub;
    const double *__restrict__ u1 = tp(o.x, chunk) + sub;
    double *__restrict__ xs = mode == 1 ? nullptr : tp(o.d, chunk) + sub;
    const D2 a = make_double2(sm.st[ST_CW0][sub], sm.st[ST_CW0][sub + 1]);
    const D2 b = make_double2(sm.st[ST_CW1][sub], sm.st[ST_CW1][sub + 1]);
    const D2 c = make_double2(sm.st[ST_CU][sub], sm.st[ST_CU][sub + 1]);
    const D2 e = make_double2(sm.st[ST_CX][sub], sm.st[ST_CX][sub + 1]);
    if (mode == 1) {  // directions only; the solution is updated together with the next iteration's
#pragma unroll 4
        for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
            const size_t ro = (size_t)row * TW;
            const D2 w0v = ld2c(w0 + ro), w1v = ld2c(w1 + ro), uv = ld2c(u1 + ro);
            st2(w0 + ro, make_double2(fma(a.x, w0v.x, fma(b.x, w1v.x, c.x * uv.x)), fma(a.y, w0v.y, fma(b.y, w1v.y, c.y * uv.y))));
        }
        return;
    }
#pragma unroll 4
    for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
        const size_t ro = (size_t)row * TW;
        const D2 w0v = ld2c(w0 + ro), w1v = ld2c(w1 + ro), uv = ld2c(u1 + ro);
        D2 xv = ld2c(xs + ro);
        const D2 wn = make_double2(fma(a.x, w0v.x, fma(b.x, w1v.x, c.x * uv.x)), fma(a.y, w0v.y, fma(b.y, w1v.y, c.y * uv.y)));
        if (mode == 2) {  // the previous iteration's update first: same operations in the same order as undeferred
            if (ep.x != 0.0) xv.x = fma(ep.x, w1v.x, xv.x);
            if (ep.y != 0.0) xv.y = fma(ep.y, w1v.y, xv.y);
        }
        if (e.x != 0.0) xv.x = fma(e.x, wn.x, xv.x);
        if (e.y != 0.0) xv.y = fma(e.y, wn.y, xv.y);
        st2(w0 + ro, wn);
        st2(xs + ro, xv);
    }
}

// one Chebyshev step of op_cheb_small for row `row`: the operator entries (and, weighted, the weight rows) are read from
// shared memory (SH: staged once per operation) or through the read-only path from L2
template <bool WEIGHTED, bool SH>
__device__ __forceinline__ D2 cheb_small_row_sum(const int *__restrict__ off, const unsigned char *__restrict__ pk, const double *__restrict__ V,
                                                 const double *zin, int row)
{
    constexpr int ES = WEIGHTED ? 16 : 12;
    const int sl = row / SLICE, rs = row % SLICE;
    const int k0 = __ldg(off + sl), w = __ldg(off + sl + 1) - k0;
    const unsigned char *base = pk + (size_t)k0 * (SLICE * ES);
    const double *eval = reinterpret_cast<const double *>(base) + rs;
    const int *ecol = reinterpret_cast<const int *>(base + (size_t)w * (SLICE * 8)) + rs;
    const int *ewid = ecol + w * SLICE;
    D2 s = make_double2(0.0, 0.0);
    for (int k = 0; k < w; ++k) {
        const double c = SH ? eval[k * SLICE] : __ldg(eval + k * SLICE);
        const D2 xv = ld2c(zin + (size_t)(SH ? ecol[k * SLICE] : __ldg(ecol + k * SLICE)) * TW);
        if (WEIGHTED) {
            const D2 wv = ld2c(V + (size_t)(SH ? ewid[k * SLICE] : __ldg(ewid + k * SLICE)) * TW);
            s.x = fma(c * wv.x, xv.x, s.x);
            s.y = fma(c * wv.y, xv.y, s.y);
        } else {
            s.x = fma(c, xv.x, s.x);
            s.y = fma(c, xv.y, s.y);
        }
    }
    return s;
}

// The coarsest level of the Schur V-cycle (and the whole Schur preconditioner of the coarsest mesh level) is a
// Chebyshev iteration of degree a0 on a few dozen rows.  As separate operations every step costs the fixed latency of
// an operation (operation fetch, first memory round trip, barrier: 4-14 k cycles for 2 KB of data).  Here the whole
// iteration is ONE operation: the iterates z, d live in shared memory (the staging buffers, idle meanwhile), the steps
// are separated by CTA barriers, r and 1/l1 are re-read from L1.  The arithmetic per row is the one of OP_CHEB_FIRST
// followed by OP_SPMM/EP_CHEB steps with the same host-computed coefficients (o.val = {ca_j, cb_j}), so the results
// are bitwise those of the unfused sequence.  With clusters / grid groups the first CTA of the tile does the work.
template <int NTt, int CS, bool WEIGHTED>
__device__ __forceinline__ void op_cheb_small(const Op &o, double *chunk, Smem &sm, unsigned char *scratch)
{
    const int sub = (threadIdx.x % LPR) * PW;
    const int n = o.n, deg = o.a0;
    const bool mine = cluster_rank<CS>(sm) == 0;
    const double *__restrict__ r = tp(o.r, chunk) + sub;
    double *__restrict__ zg = tp(o.y, chunk) + sub;
    const double *V = WEIGHTED ? tp(o.v, chunk) + sub : nullptr;
    const double *__restrict__ dinvb = WEIGHTED ? tp(o.w, chunk) + sub : nullptr;
    const double *__restrict__ coef = o.val;
    double *zs0 = reinterpret_cast<double *>(scratch) + sub, *zs1 = zs0 + SMALLN * TW, *ds = zs1 + SMALLN * TW;
    constexpr int ES = WEIGHTED ? 16 : 12;
    // The operator's entries and the tile's weight rows are copied into shared memory once (when the CTA's staging area has
    // room behind the iterates): every step then reads them at shared-memory latency instead of paying two dependent L2
    // round trips (entries -> weights).  o.a1 = number of weight rows (weighted operators).
    constexpr size_t AVAIL = (size_t)(NTt / 32) * NSTAGE * sizeof(WarpStage);
    constexpr size_t ITER_BYTES = 3 * (size_t)SMALLN * TW * sizeof(double);
    const int nsl = (n + SLICE - 1) / SLICE;
    const size_t ebytes = (size_t)__ldg(o.rowptr + nsl) * (SLICE * ES);
    const size_t vbytes = WEIGHTED ? (size_t)o.a1 * TW * sizeof(double) : 0;
    const bool staged = deg > 1 && AVAIL > ITER_BYTES && ebytes + vbytes <= AVAIL - ITER_BYTES && (!WEIGHTED || o.a1 > 0);
    const unsigned char *pk = o.pk;
    if (mine && staged) {
        unsigned char *se = scratch + ITER_BYTES;
        const uint4 *src = reinterpret_cast<const uint4 *>(o.pk);
        uint4 *dst = reinterpret_cast<uint4 *>(se);
        const int ne16 = (int)(ebytes / 16);
        for (int i = threadIdx.x; i < ne16; i += NTt) dst[i] = __ldg(src + i);
        pk = se;
        if (WEIGHTED) {
            double *sv = reinterpret_cast<double *>(se + ebytes);
            const double2 *vsrc = reinterpret_cast<const double2 *>(tp(o.v, chunk));
            double2 *vdst = reinterpret_cast<double2 *>(sv);
            const int nv16 = (int)(vbytes / 16);
            for (int i = threadIdx.x; i < nv16; i += NTt) vdst[i] = vsrc[i];
            V = sv + sub;
        }
        // the first step (j = 0) reads neither; the barrier that ends it publishes the copies
    }
    D2 acc = make_double2(0.0, 0.0);
    if (mine) {
        double *zin = zs0, *zout = zs1;
        for (int j = 0; j < deg; ++j) {
            const double ca = __ldg(coef + 2 * j), cb = __ldg(coef + 2 * j + 1);
            for (int row = threadIdx.x / LPR; row < n; row += NTt / LPR) {
                const size_t ro = (size_t)row * TW;
                const D2 rv = ld2c(r + ro);
                D2 di;
                if (WEIGHTED) di = ld2c(dinvb + ro);
                else { const double t = __ldg(o.fixed + row); di = make_double2(t, t); }
                D2 dn, zn;
                if (j == 0) {  // OP_CHEB_FIRST: d = z = cb dinv r
                    dn = make_double2(cb * di.x * rv.x, cb * di.y * rv.y);
                    zn = dn;
                } else {       // EP_CHEB: d = ca d + cb dinv (r - A z);  z' = z + d
                    const D2 s = staged ? cheb_small_row_sum<WEIGHTED, true>(o.rowptr, pk, V, zin, row)
                                        : cheb_small_row_sum<WEIGHTED, false>(o.rowptr, pk, V, zin, row);
                    dn = make_double2(cb * di.x * (rv.x - s.x), cb * di.y * (rv.y - s.y));
                    if (ca != 0.0) {
                        const D2 dv = ld2c(ds + ro);
                        dn.x = fma(ca, dv.x, dn.x);
                        dn.y = fma(ca, dv.y, dn.y);
                    }
                    const D2 zv = ld2c(zin + ro);
                    zn = make_double2(zv.x + dn.x, zv.y + dn.y);
                }
                st2(ds + ro, dn);
                st2(zout + ro, zn);
                if (j == deg - 1) {
                    st2(zg + ro, zn);
                    acc.x = fma(zn.x, rv.x, acc.x);
                    acc.y = fma(zn.y, rv.y, acc.y);
                }
            }
            __syncthreads();
            double *t = zin; zin = zout; zout = t;
        }
    }
    if (o.flags & F_DOT) block_dot<NTt, CS>(acc, sm, o.slot, (o.flags & F_DOT_ACC) != 0);
}

template <int NTt, int CS>
__device__ __forceinline__ void op_setup_spmm(const Op &o, double *chunk, const Smem &sm)
{
    const int sub = (threadIdx.x % LPR) * PW;
    int r0, r1;
    my_rows<CS>(o, sm, r0, r1);
    const double *__restrict__ x = tp(o.x, chunk) + sub;
    double *__restrict__ y = tp(o.y, chunk) + sub;
    const bool absx = (o.flags & F_ABSX) != 0, recip = (o.flags & F_RECIP) != 0;
    for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
        D2 s = make_double2(0.0, 0.0);
        const int p0 = __ldg(o.rowptr + row), p1 = __ldg(o.rowptr + row + 1);
#pragma unroll 4
        for (int p = p0; p < p1; ++p) {
            const double c = __ldg(o.val + p);
            D2 xv = ld2c(x + (size_t)__ldg(o.col + p) * TW);
            if (absx) { xv.x = fabs(xv.x); xv.y = fabs(xv.y); }
            s.x = fma(c, xv.x, s.x);
            s.y = fma(c, xv.y, s.y);
        }
        if (recip) { s.x = safe_inv(s.x); s.y = safe_inv(s.y); }
        st2(y + (size_t)row * TW, s);
    }
}

// ---- per-sample scalar recurrences of preconditioned MINRES (mfem::MINRESSolver::Mult structure, reached by the
// reference through ParELAG's Krylov wrapper at /root/reference/src/PDESampler.cpp:517-522 and
// src/DarcySolver.cpp:629-631), with lazily normalised Lanczos vectors.  Thread j < TW owns sample j of the tile.
__device__ __forceinline__ void sc_init(const Op &o, int tile, Smem &sm, const ProgParams &P)
{
    const int j = threadIdx.x;
    if (j >= TW) return;
    const double eta = sqrt(fmax(sm.dots[o.slot][j], 0.0));
    const double goal = fmax(P.rel * eta, P.abs_);
    const int sample = tile * TW + j;
    sm.st[ST_BETA][j] = eta;
    sm.st[ST_IB][j] = safe_inv(eta);
    sm.st[ST_IBPREV][j] = 0.0;
    sm.st[ST_G0][j] = 1.0;
    sm.st[ST_G1][j] = 1.0;
    sm.st[ST_S0][j] = 0.0;
    sm.st[ST_S1][j] = 0.0;
    sm.st[ST_ETA][j] = eta;
    sm.st[ST_GOAL][j] = goal;
    sm.st[ST_CXP][j] = 0.0;
    sm.st[ST_OM0][j] = 0.0;
    sm.st[ST_OM1][j] = 0.0;
    sm.st[ST_QACC][j] = 0.0;
    sm.active[j] = (sample < P.nsamples) && (eta > goal);
    sm.iters[j] = 0;
}

__device__ __forceinline__ void sc_alpha(const Op &o, Smem &sm)
{
    const int j = threadIdx.x;
    if (j >= TW) return;
    double cq = 0.0, cv1 = 0.0, cv0 = 0.0;
    if (sm.active[j]) {
        const double ib = sm.st[ST_IB][j];
        const double alpha = sm.dots[o.slot][j] * ib * ib;
        sm.st[ST_ALPHA][j] = alpha;
        cq = ib;
        cv1 = -alpha * ib;
        cv0 = -sm.st[ST_BETA][j] * sm.st[ST_IBPREV][j];
    }
    sm.st[ST_CQ][j] = cq;
    sm.st[ST_CV1][j] = cv1;
    sm.st[ST_CV0][j] = cv0;
}

__device__ __forceinline__ void sc_beta(const Op &o, Smem &sm, const ProgParams &P)
{
    const int j = threadIdx.x;
    if (j >= TW) return;
    double cw0 = 0.0, cw1 = 0.0, cu = 0.0, cx = 0.0;
    if (sm.active[j]) {
        const double beta_new = sqrt(fmax(sm.dots[o.slot][j], 0.0));
        const double beta = sm.st[ST_BETA][j], alpha = sm.st[ST_ALPHA][j], ib = sm.st[ST_IB][j];
        double g0 = sm.st[ST_G0][j], g1 = sm.st[ST_G1][j], s0 = sm.st[ST_S0][j], s1 = sm.st[ST_S1][j];
        double eta = sm.st[ST_ETA][j];
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
        const int it = sm.iters[j] + 1;
        sm.iters[j] = it;
        sm.st[ST_G0][j] = g0;
        sm.st[ST_G1][j] = g1;
        sm.st[ST_S0][j] = s0;
        sm.st[ST_S1][j] = s1;
        sm.st[ST_ETA][j] = eta;
        sm.st[ST_IBPREV][j] = ib;
        sm.st[ST_BETA][j] = beta_new;
        sm.st[ST_IB][j] = safe_inv(beta_new);
        if (o.a1) {
            // Only a linear functional Q = obs . x of the solution is wanted (DarcySolver::SolveFwd returns Q and C, not the
            // solution): x = sum_k cx_k w_k with w_k = cw0 w_{k-2} + cw1 w_{k-1} + cu z_k, so obs . w_k obeys the same recurrence
            // driven by the scalars obs . z_k (dots[3], an OP_DOT_FIXED earlier in this iteration) and the direction vectors
            // and the solution vector are never formed.
            const double om = fma(cw0, sm.st[ST_OM0][j], fma(cw1, sm.st[ST_OM1][j], cu * sm.dots[3][j]));
            sm.st[ST_OM0][j] = sm.st[ST_OM1][j];
            sm.st[ST_OM1][j] = om;
            sm.st[ST_QACC][j] = fma(cx, om, sm.st[ST_QACC][j]);
        }
        if (fabs(eta) <= sm.st[ST_GOAL][j] || it >= P.max_iter || beta_new == 0.0) sm.active[j] = 0;
    }
    sm.st[ST_CW0][j] = cw0;
    sm.st[ST_CW1][j] = cw1;
    sm.st[ST_CU][j] = cu;
    sm.st[ST_CX][j] = cx;
    if (o.a0) sm.st[ST_CXP][j] = cx;  // this iteration's solution update is deferred to the next one
}

// ---- Jacobi-preconditioned CG (sampler, SPD form): per-sample scalars; thread j < TW owns sample j of the tile ----------
__device__ __forceinline__ void cg_init(const Op &o, int tile, Smem &sm, const ProgParams &P)
{
    const int j = threadIdx.x;
    if (j >= TW) return;
    const double rz = fmax(sm.dots[o.slot][j], 0.0);
    const double eta = sqrt(rz);
    const double goal = fmax(P.rel * eta, P.abs_);
    sm.st[ST_RZ][j] = rz;
    sm.st[ST_GOAL][j] = goal;
    sm.st[ST_CGA][j] = 0.0;
    sm.st[ST_CGB][j] = 0.0;
    sm.active[j] = (tile * TW + j < P.nsamples) && (eta > goal);
    sm.iters[j] = 0;
}
__device__ __forceinline__ void cg_alpha(const Op &o, Smem &sm)
{
    const int j = threadIdx.x;
    if (j >= TW) return;
    const double pq = sm.dots[o.slot][j];
    double a = 0.0;
    if (sm.active[j]) {
        if (pq > 0.0) a = sm.st[ST_RZ][j] / pq;
        else sm.active[j] = 0;   // breakdown (p = 0): nothing left to correct
    }
    sm.st[ST_CGA][j] = a;
}
__device__ __forceinline__ void cg_beta(const Op &o, Smem &sm, const ProgParams &P)
{
    const int j = threadIdx.x;
    if (j >= TW) return;
    double b = 0.0;
    if (sm.active[j]) {
        const double rzn = fmax(sm.dots[o.slot][j], 0.0), rz = sm.st[ST_RZ][j];
        b = rz > 0.0 ? rzn / rz : 0.0;
        sm.st[ST_RZ][j] = rzn;
        const int it = sm.iters[j] + 1;
        sm.iters[j] = it;
        if (sqrt(rzn) <= sm.st[ST_GOAL][j] || it >= P.max_iter) sm.active[j] = 0;
    }
    sm.st[ST_CGB][j] = b;
}

// ---- Chebyshev semi-iteration (sampler, SPD form): the step count for the requested reduction is known a priori from the
// spectrum of the (sample-independent) operator; the true residual norm is checked after every block of steps ----------
__device__ __forceinline__ void chb_init(const Op &o, int tile, Smem &sm, const ProgParams &P)
{
    const int j = threadIdx.x;
    if (j >= TW) return;
    const double eta = sqrt(fmax(o.ca * sm.dots[o.slot][j], 0.0));
    const double goal = fmax(P.rel * eta, P.abs_);
    sm.st[ST_GOAL][j] = goal;
    sm.active[j] = (tile * TW + j < P.nsamples) && (eta > goal);
    sm.iters[j] = 0;
}
__device__ __forceinline__ void chb_check(const Op &o, Smem &sm, const ProgParams &P)
{
    const int j = threadIdx.x;
    if (j >= TW) return;
    if (sm.active[j]) {
        const int it = sm.iters[j] + o.a0;
        sm.iters[j] = it;
        if (sqrt(fmax(sm.dots[o.slot][j], 0.0)) <= sm.st[ST_GOAL][j] || it >= P.max_iter) sm.active[j] = 0;
    }
}

// x += a p ; r -= a q ; rz' = sum dinv r^2   (a = 0 for converged samples: their vectors stay as they are)
template <int NTt, int CS>
__device__ __forceinline__ void op_cg_update(const Op &o, double *chunk, Smem &sm)
{
    const int sub = (threadIdx.x % LPR) * PW;
    int r0, r1;
    my_rows<CS>(o, sm, r0, r1);
    const double *__restrict__ p = tp(o.x, chunk) + sub;
    const double *__restrict__ q = tp(o.r, chunk) + sub;
    double *__restrict__ res = tp(o.y, chunk) + sub;
    double *__restrict__ xs = tp(o.d, chunk) + sub;
    const D2 a = make_double2(sm.st[ST_CGA][sub], sm.st[ST_CGA][sub + 1]);
    D2 acc = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
        const size_t ro = (size_t)row * TW;
        const D2 pv = ld2c(p + ro), qv = ld2c(q + ro);
        D2 rv = ld2c(res + ro), xv = ld2c(xs + ro);
        const double di = __ldg(o.fixed + row);
        xv.x = fma(a.x, pv.x, xv.x);
        xv.y = fma(a.y, pv.y, xv.y);
        rv.x = fma(-a.x, qv.x, rv.x);
        rv.y = fma(-a.y, qv.y, rv.y);
        st2(xs + ro, xv);
        st2(res + ro, rv);
        acc.x = fma(di * rv.x, rv.x, acc.x);
        acc.y = fma(di * rv.y, rv.y, acc.y);
    }
    block_dot<NTt, CS>(acc, sm, o.slot, false);
}

// p = dinv r + b p
template <int NTt, int CS>
__device__ __forceinline__ void op_cg_dir(const Op &o, double *chunk, const Smem &sm)
{
    const int sub = (threadIdx.x % LPR) * PW;
    int r0, r1;
    my_rows<CS>(o, sm, r0, r1);
    const double *__restrict__ res = tp(o.x, chunk) + sub;
    double *__restrict__ p = tp(o.y, chunk) + sub;
    const D2 b = make_double2(sm.st[ST_CGB][sub], sm.st[ST_CGB][sub + 1]);
#pragma unroll 4
    for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
        const size_t ro = (size_t)row * TW;
        const D2 rv = ld2c(res + ro), pv = ld2c(p + ro);
        const double di = __ldg(o.fixed + row);
        st2(p + ro, make_double2(fma(b.x, pv.x, di * rv.x), fma(b.y, pv.y, di * rv.y)));
    }
}

// Noise generation fused with the SPDE right-hand-side scaling: thread (chunk c, sample j) jumps to stream position
// u0 + sample * n + c * T and steps T times (PDESampler::Sample + :352-358 of src/PDESampler.cpp).
template <int NTt, int CS>
__device__ __forceinline__ void op_rng(const Op &o, int tile, double *chunk, const ProgParams &P, const Smem &sm)
{
    double *__restrict__ y = tp(o.y, chunk);
    const int j = threadIdx.x & (TW - 1), c = cluster_rank<CS>(sm) * (NTt / TW) + threadIdx.x / TW;
    const int NCH = (CS == 0 ? sm.grp_size : CS) * NTt / TW;
    const int T = (o.n + NCH - 1) / NCH;
    const int i0 = c * T, i1 = min(o.n, i0 + T);
    const int sample = tile * TW + j;
    if (i0 >= i1) return;
    if (sample >= P.nsamples) {
        for (int i = i0; i < i1; ++i) y[(size_t)i * TW + j] = 0.0;
        return;
    }
    uint32_t r[5], co[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { r[k] = P.tab->r0[k]; co[k] = P.tab->a[k]; }
    yarn5_jump(r, o.u0 + (uint64_t)sample * (uint64_t)o.a0 * (uint64_t)o.n + (uint64_t)i0, P.tab->jump);
    for (int i = i0; i < i1; ++i) {
        yarn5_step(r, co);
        const uint32_t v = yarn5_output(r[0], P.tab->powtab);
        const double z = dev_normal_from_engine(v, P.mu, P.sigma);
        y[(size_t)i * TW + j] = dm(dm(o.ca, z), __ldg(o.fixed + i));
    }
}

template <int NTt, int MINB, int CS>
__global__ void __launch_bounds__(NTt, MINB) k_run_program(const ProgParams P)
{
    __shared__ Smem sm;
    extern __shared__ __align__(128) unsigned char dyn_smem[];  // NTt/32 warps x NSTAGE staging buffers
    const int gsz = CS == 0 ? P.group : CS;
    const int tile = blockIdx.x / gsz;  // the CTAs of a cluster / grid group share a tile
    if (CS == 0 && threadIdx.x == 0) {
        sm.grp_size = gsz;
        sm.grp_rank = blockIdx.x % gsz;
        sm.grp_bar = P.grp_bar + 2 * (size_t)tile;
        sm.grp_part = P.grp_part + (size_t)tile * gsz * TW;
    }
    double *const chunk = P.base + (size_t)tile * (size_t)P.chunk;
    if (threadIdx.x < KC_COUNT) { sm.cyc[threadIdx.x] = 0ull; sm.cbytes[threadIdx.x] = 0.0; sm.cops[threadIdx.x] = 0u; }
    if (threadIdx.x < TW) { sm.active[threadIdx.x] = 0; sm.iters[threadIdx.x] = 0; }
    StageCtx sc;
    sc.buf = reinterpret_cast<WarpStage *>(dyn_smem) + (threadIdx.x >> 5) * NSTAGE;
    sc.bar = smem_u32(&sm.bars[threadIdx.x >> 5][0]);
    sc.phase = 0u;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < NSTAGE; ++j) mbar_init(sc.bar + 8 * j, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int crank = cluster_rank<CS>(sm);
    const long long t_begin = clock64();
    int pc = 0;
    while (pc < P.nops) {
        const Op &o = P.ops[pc];
        const int kind = o.kind, flags = o.flags;
        int next = pc + 1;
        const long long t0 = clock64();
        switch (kind) {
        case OP_SPMM: {
            const int ep = (flags >> F_EP_SHIFT) & 3;
            const bool w = flags & F_WEIGHTED, dot = flags & F_DOT;
#define PMC_SPMM(EP_, W_, BD_, DOT_)                                                          \
    do {                                                                                      \
        if (flags & F_STAGED) op_spmm<NTt, CS, EP_, W_, BD_, DOT_, true>(o, chunk, sm, sc);   \
        else if (flags & F_CHUNKED) op_spmm_chunked<NTt, CS, EP_, W_, BD_, DOT_>(o, chunk, sm, sc); \
        else op_spmm<NTt, CS, EP_, W_, BD_, DOT_, false>(o, chunk, sm, sc);                   \
    } while (0)
            if (ep == EP_AX) {
                if (w) { if (dot) PMC_SPMM(EP_AX, true, false, true); else PMC_SPMM(EP_AX, true, false, false); }
                else   { if (dot) PMC_SPMM(EP_AX, false, false, true); else PMC_SPMM(EP_AX, false, false, false); }
            } else if (ep == EP_RESID) {
                if (w) PMC_SPMM(EP_RESID, true, false, false); else PMC_SPMM(EP_RESID, false, false, false);
            } else if (ep == EP_ADD) {
                PMC_SPMM(EP_ADD, false, false, false);
            } else {
                if (w) { if (dot) PMC_SPMM(EP_CHEB, true, true, true); else PMC_SPMM(EP_CHEB, true, true, false); }
                else if (flags & F_THREE) {
                    if (flags & F_STAGED) op_spmm<NTt, CS, EP_CHEB, false, false, false, true, true>(o, chunk, sm, sc);
                    else if (flags & F_CHUNKED) op_spmm_chunked<NTt, CS, EP_CHEB, false, false, false, true>(o, chunk, sm, sc);
                    else op_spmm<NTt, CS, EP_CHEB, false, false, false, false, true>(o, chunk, sm, sc);
                }
                else   { if (dot) PMC_SPMM(EP_CHEB, false, false, true); else PMC_SPMM(EP_CHEB, false, false, false); }
            }
#undef PMC_SPMM
        } break;
        case OP_CHEB_FIRST:
            if (flags & F_BDINV) op_cheb_first<NTt, CS, true>(o, chunk, sm); else op_cheb_first<NTt, CS, false>(o, chunk, sm);
            break;
        case OP_LINCOMB3:
            if (flags & F_DOT) { if (flags & F_BDINV) op_lincomb3<NTt, CS, true, true>(o, chunk, sm); else op_lincomb3<NTt, CS, true, false>(o, chunk, sm); }
            else op_lincomb3<NTt, CS, false, false>(o, chunk, sm);
            break;
        case OP_SOL_UPDATE: op_sol_update<NTt, CS>(o, chunk, sm); break;
        case OP_SETUP_SPMM: op_setup_spmm<NTt, CS>(o, chunk, sm); break;
        case OP_FILL: {
            double *y = tp(o.y, chunk);
            const double v = o.ca;
            int r0, r1;
            my_rows<CS>(o, sm, r0, r1);
            for (int i = r0 * (TW / 2) + threadIdx.x; i < r1 * (TW / 2); i += NTt) reinterpret_cast<double2 *>(y)[i] = make_double2(v, v);
        } break;
        case OP_COPY: {
            const double2 *x = reinterpret_cast<const double2 *>(tp(o.x, chunk));
            double2 *y = reinterpret_cast<double2 *>(tp(o.y, chunk));
            int r0, r1;
            my_rows<CS>(o, sm, r0, r1);
#pragma unroll 4
            for (int i = r0 * (TW / 2) + threadIdx.x; i < r1 * (TW / 2); i += NTt) y[i] = x[i];
        } break;
        case OP_BROADCAST: {
            double *y = tp(o.y, chunk);
            const int sub = (threadIdx.x % LPR) * PW;
            int r0, r1;
            my_rows<CS>(o, sm, r0, r1);
            for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
                const double v = __ldg(o.fixed + row);
                st2(y + (size_t)row * TW + sub, make_double2((tile * TW + sub < P.nsamples) ? v : 0.0,
                                                              (tile * TW + sub + 1 < P.nsamples) ? v : 0.0));
            }
        } break;
        case OP_MAP_EXP: {
            const double *x = tp(o.x, chunk);
            double *y = tp(o.y, chunk);
            int r0, r1;
            my_rows<CS>(o, sm, r0, r1);
            for (int i = r0 * TW + threadIdx.x; i < r1 * TW; i += NTt) y[i] = exp(x[i]);
        } break;
        case OP_DOT_FIXED: {
            const int sub = (threadIdx.x % LPR) * PW;
            const double *x = tp(o.x, chunk) + sub;
            D2 acc = make_double2(0.0, 0.0);
            int r0, r1;
            my_rows<CS>(o, sm, r0, r1);
            for (int row = r0 + threadIdx.x / LPR; row < r1; row += NTt / LPR) {
                const double w = __ldg(o.fixed + row);
                if (w != 0.0) {
                    const D2 xv = ld2c(x + (size_t)row * TW);
                    acc.x = fma(w, xv.x, acc.x);
                    acc.y = fma(w, xv.y, acc.y);
                }
            }
            block_dot<NTt, CS>(acc, sm, 3, false);
            __syncthreads();
            if (threadIdx.x < TW && crank == 0 && o.y.off >= 0) tp(o.y, chunk)[threadIdx.x] = sm.dots[3][threadIdx.x];
        } break;
        case OP_DOT_SPARSE: {
            const int sub = (threadIdx.x % LPR) * PW;
            const double *x = tp(o.x, chunk) + sub;
            D2 acc = make_double2(0.0, 0.0);
            int r0, r1;
            my_rows<CS>(o, sm, r0, r1);
            for (int i = r0 + threadIdx.x / LPR; i < r1; i += NTt / LPR) {
                const double w = __ldg(o.val + i);
                const D2 xv = ld2c(x + (size_t)__ldg(o.col + i) * TW);
                acc.x = fma(w, xv.x, acc.x);
                acc.y = fma(w, xv.y, acc.y);
            }
            block_dot<NTt, CS>(acc, sm, 3, false);
        } break;
        case OP_SC_INIT: sc_init(o, tile, sm, P); break;
        case OP_SC_ALPHA: sc_alpha(o, sm); break;
        case OP_SC_BETA: sc_beta(o, sm, P); break;
        case OP_CHECK: {
            int any = 0;
#pragma unroll
            for (int j = 0; j < TW; ++j) any |= sm.active[j];
            if (!any) next = o.a0;
        } break;
        case OP_JUMP: next = o.a0; break;
        case OP_STORE_ITERS:
            if (threadIdx.x < TW && crank == 0) {
                const int sample = tile * TW + threadIdx.x;
                if (o.y.off >= 0) tp(o.y, chunk)[threadIdx.x] = (double)sm.iters[threadIdx.x];
                if (sample < P.nsamples) atomicAdd(&P.stats->iters_total, (unsigned long long)sm.iters[threadIdx.x]);
            }
            break;
        case OP_RNG: op_rng<NTt, CS>(o, tile, chunk, P, sm); break;
        case OP_CHEB_SMALL:
            if (flags & F_WEIGHTED) op_cheb_small<NTt, CS, true>(o, chunk, sm, dyn_smem);
            else op_cheb_small<NTt, CS, false>(o, chunk, sm, dyn_smem);
            break;
        case OP_CG_INIT: cg_init(o, tile, sm, P); break;
        case OP_CG_ALPHA: cg_alpha(o, sm); break;
        case OP_CG_BETA: cg_beta(o, sm, P); break;
        case OP_CHB_INIT: chb_init(o, tile, sm, P); break;
        case OP_CHB_CHECK: chb_check(o, sm, P); break;
        case OP_STORE_Q:
            if (threadIdx.x < TW && crank == 0) tp(o.y, chunk)[threadIdx.x] = sm.st[ST_QACC][threadIdx.x];
            break;
        case OP_CG_UPDATE: op_cg_update<NTt, CS>(o, chunk, sm); break;
        case OP_CG_DIR: op_cg_dir<NTt, CS>(o, chunk, sm); break;
        case OP_LIKELIHOOD:
            // BayesianInverseProblem::ComputeLikelihood / ComputeR (/root/reference/src/BayesianInverseProblem.cpp:190-218)
            if (threadIdx.x < TW && crank == 0) {
                const double *G = tp(o.x, chunk);
                double acc2 = 0.0;
                for (int i = 0; i < o.n; ++i) {
                    const double dlt = G[(size_t)i * TW + threadIdx.x] - __ldg(o.fixed + i);
                    acc2 = fma(dlt, dlt, acc2);
                }
                double like = exp(-acc2 * o.ca);
                if (o.r.off >= 0) like *= tp(o.r, chunk)[threadIdx.x];
                tp(o.y, chunk)[threadIdx.x] = like;
            }
            break;
        default: break;
        }
        op_barrier<CS>(sm, flags);
        if (threadIdx.x == 0) {
            if (P.op_cycles) {
                atomicAdd(&P.op_cycles[2 * pc], (unsigned long long)(clock64() - t0));
                atomicAdd(&P.op_cycles[2 * pc + 1], 1ull);
            }
            sm.cyc[o.kclass] += (unsigned long long)(clock64() - t0);
            if (crank == 0) {  // bytes and operation counts once per tile
                double b = o.bytes;
                if (flags & F_INLOOP) {   // rows are 4 realisations wide: the lanes of converged ones move, but are not credited
                    int na = 0;
#pragma unroll
                    for (int j = 0; j < TW; ++j) na += sm.active[j] ? 1 : 0;
                    b *= 0.25 * (double)na;
                }
                sm.cbytes[o.kclass] += b;
                sm.cops[o.kclass] += 1u;
            }
        }
        pc = next;
    }
    op_barrier<CS>(sm, 0);  // no CTA of a cluster exits while another may still read its shared memory
    if (threadIdx.x < KC_COUNT && sm.cyc[threadIdx.x]) atomicAdd(&P.stats->class_cycles[threadIdx.x], sm.cyc[threadIdx.x]);
    if (threadIdx.x < KC_COUNT && sm.cops[threadIdx.x]) {
        atomicAdd(&P.stats->class_bytes[threadIdx.x], sm.cbytes[threadIdx.x]);
        atomicAdd(&P.stats->class_ops[threadIdx.x], (unsigned long long)sm.cops[threadIdx.x]);
        atomicAdd(&P.stats->bytes, sm.cbytes[threadIdx.x]);
        atomicAdd(&P.stats->ops_executed, (unsigned long long)sm.cops[threadIdx.x]);
    }
    if (threadIdx.x == 0) atomicAdd(&P.stats->cta_cycles, (unsigned long long)(clock64() - t_begin));
}

// ---- layout conversion and small reductions (separate launches; not on the per-iteration path) ----------------
// Host layout [nsamples][n] (sample-major) -> rows of the tile chunks (dst = base + operand offset);  MODE 1 applies the SPDE right-hand-side scaling
// out = (-g * xi) * w_sqrt[row] (/root/reference/src/PDESampler.cpp:352-358).
template <int MODE>
__global__ void k_to_tiles(int n, long long chunk, int ntiles, int nsamples, const double *__restrict__ src,
                           double *__restrict__ dst, double neg_g, const double *__restrict__ w_sqrt,
                           const int *__restrict__ rowmap)
{
    const size_t total = (size_t)ntiles * n * TW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % TW);
        const size_t t = i / TW;
        const int row = (int)(t % n), tile = (int)(t / n);
        const int s = tile * TW + j;
        double v = s < nsamples ? src[(size_t)s * n + row] : 0.0;
        if (MODE == 1) v = __dmul_rn(__dmul_rn(neg_g, v), __ldg(w_sqrt + row));
        const int irow = rowmap ? __ldg(rowmap + row) : row;  // the library's internal numbering of the row
        dst[(size_t)tile * (size_t)chunk + (size_t)irow * TW + j] = v;
    }
}

// rows of the tile chunks (src = base + operand offset) -> host layout [nsamples][n];  MODE 1 applies exp.
template <int MODE>
__global__ void k_from_tiles(int n, long long chunk, int nsamples, const double *__restrict__ src, double *__restrict__ dst,
                             const int *__restrict__ rowmap)
{
    const size_t total = (size_t)nsamples * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i % n), s = (int)(i / n);
        const int tile = s / TW, j = s % TW;
        const int irow = rowmap ? __ldg(rowmap + row) : row;
        double v = src[(size_t)tile * (size_t)chunk + (size_t)irow * TW + j];
        if (MODE == 1) v = exp(v);
        dst[i] = v;
    }
}

__global__ void k_fill(double *p, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// Per-level moment sums of MLMC_Manager::InitRun (/root/reference/src/MLMC_Manager.cpp:123-131,:158-168), order
// {Y2, Y, ABSY, Q2, Q, ABSQ, C, Y3, Y4}.  One CTA, fixed reduction tree => deterministic.
__global__ void __launch_bounds__(256)
    k_mlmc_accumulate(int nsamples, const double *__restrict__ base, long long chunk, long long offQ, long long offQc,
                      double cost, double *__restrict__ out9, double *__restrict__ rows)
{
    __shared__ double red[9][256];
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = threadIdx.x; j < nsamples; j += 256) {
        const size_t to = (size_t)(j / TW) * (size_t)chunk + (size_t)(j % TW);
        const double q = base[to + offQ], qc = offQc >= 0 ? base[to + offQc] : 0.0;
        const double y = offQc >= 0 ? q - qc : q;
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

// Sums of ML_BayesRatio_Manager::InitRun (/root/reference/src/ML_BayesRatio_Manager.hpp:313-424) in the layout of its
// enum (:67-70): {YZ2, YZ, ABS_YZ, Z2, Z, ABS_Z, YR2, YR, ABS_YR, R2, R, ABS_R, ..., C = 18}.  One CTA, fixed tree.
__global__ void __launch_bounds__(256)
    k_bayes_accumulate(int nsamples, const double *__restrict__ base, long long chunk, long long off_r, long long off_rc,
                       long long off_z, long long off_zc, double cost, double *__restrict__ out20, double *__restrict__ rows)
{
    __shared__ double red[13][256];
    double a[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = threadIdx.x; j < nsamples; j += 256) {
        const size_t to = (size_t)(j / TW) * (size_t)chunk + (size_t)(j % TW);
        const double r = base[to + off_r], z = base[to + off_z];
        const double yr = off_rc >= 0 ? r - base[to + off_rc] : r;
        const double yz = off_zc >= 0 ? z - base[to + off_zc] : z;
        a[0] += yz * yz; a[1] += yz; a[2] += fabs(yz); a[3] += z * z; a[4] += z; a[5] += fabs(z);
        a[6] += yr * yr; a[7] += yr; a[8] += fabs(yr); a[9] += r * r; a[10] += r; a[11] += fabs(r);
        a[12] += cost;
        if (rows) {
            rows[5 * (size_t)j + 0] = r;
            rows[5 * (size_t)j + 1] = yr;
            rows[5 * (size_t)j + 2] = z;
            rows[5 * (size_t)j + 3] = yz;
            rows[5 * (size_t)j + 4] = cost;
        }
    }
    for (int k = 0; k < 13; ++k) red[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int k = 0; k < 13; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 12) out20[threadIdx.x] = red[threadIdx.x][0];
    if (threadIdx.x == 12) out20[18] = red[12][0];
}

}  // namespace pmc
